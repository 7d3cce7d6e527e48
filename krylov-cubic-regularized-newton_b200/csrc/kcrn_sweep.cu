// K1 / K3 / K4: the sweeps over the sparse data matrix.
//
// One kernel template serves both storage orders.  The matrix is a list of segments (rows of the CSR copy,
// columns of the CSC copy).  A host-built, static tile list gives every CTA (256 threads) the same amount of
// work whatever the segment-length distribution is (power-law columns put a third of the non-zeros of an
// rcv1-shaped matrix into 100 columns):
//
//   stream tile : up to 2048 non-zeros made of whole short segments (each <= 512 entries, <= 1024 segments)
//   single tile : one segment of 513 .. 2048 entries
//   piece       : 2048 consecutive entries of a segment longer than 2048
//
// One CTA per tile (the hardware scheduler balances the tiles; nothing depends on which SM runs what):
//   1. 8 (11, 12) coalesced (index, value) loads per thread, the tile's segment pointers (cp.async straight to
//      shared memory) and the epilogue operands are all requested before anything is consumed;
//   2. the dense vector is gathered through the read-only path and multiplied.  The gathers bound the sweep (one
//      L1TEX pass per distinct 128-byte line of a warp load), so the sweeps read a copy of the tile whose entries
//      are sorted by gather index -- each word carries the entry's natural position, and the product (or, with
//      LDGSTS gathers, the gathered value) goes back to that slot in shared memory: GatherVariant below;
//   3. stream tile: products -> shared memory; groups of 2^g threads (g chosen per tile on the host) sum whole
//      segments, left to right; then one thread per segment applies the epilogue, coalesced.
//      single / piece: fixed-order CTA reduction; the piece sums of a long segment are added in index order by
//      the CTA that arrives last (self-resetting ticket) -- no floating-point atomics anywhere;
//   4. the epilogue of the pass runs in registers (sigmoid weights, 1/n + l2, the Lanczos three-term update and
//      its dot product), so no extra pass over n- or d-vectors is needed.
// Every sum has a fixed order: results are run-to-run bit-identical.
//
// Warp-sized tiles were tried first (profiles/r1a_*): a lone warp walks a ~6 000-cycle dependent chain per 256
// non-zeros and an SM only holds 16-32 of them; CTA-wide tiles put 8x more non-zeros behind each chain and
// spread the per-segment work over 256 threads.
//
// Reference arithmetic being replaced: optimizer/loss.py:266-277 (A x), :215-232 (value, gradient),
// :289-302 (Hessian-vector product) and the v / w bookkeeping of optimizer/cubic.py:92-97.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "kcrn_device.cuh"
#include "kcrn_internal.cuh"

// timing probe only (profiles/microbench/build_variants.sh -DKCRN_PROBE_NOSCATTER): products stay at their gather-order
// position -- wrong sums, conflict-free shared-memory writes: what the bank conflicts of the scatter cost
#ifdef KCRN_PROBE_NOSCATTER
#define KCRN_SLOT(slot, pos) (pos)
#else
#define KCRN_SLOT(slot, pos) (slot)
#endif

// timing probe only (-DKCRN_SWEEP_TRACE, profiles/microbench/build_variants.sh): thread 0 of every CTA of the Lanczos
// passes leaves %globaltimer at the phase boundaries; kcrn_destroy prints where a pass's time goes
#ifdef KCRN_SWEEP_TRACE
#define KTRACE(k)                                                                                                   \
    do {                                                                                                            \
        if (P.trace && threadIdx.x == 0) {                                                                          \
            unsigned long long t_;                                                                                  \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                                  \
            P.trace[((size_t)P.trace_slot * 1024 + blockIdx.x) * 12 + (k)] = t_;                                    \
        }                                                                                                           \
    } while (0)
#else
#define KTRACE(k)
#endif

namespace kcrn {

// How the gathered values of a tile travel (template parameter GATHER of the sweep kernel; all three give the same
// bits, which one is fastest depends on the matrix and on the pass, so it is MEASURED per matrix: tune_* below):
//   kGatherStored   entries as stored, LDG into registers
//   kGatherOrdered  entries in gather order (fewer distinct 128-byte lines per warp load), LDG into registers,
//                   products scattered to their natural slots in shared memory
//   kGatherAsync    entries in gather order, LDGSTS (8-byte cp.async) straight into the natural slot: no registers
//                   are held while the gathers are in flight (the Lanczos passes, which carry the most state,
//                   then compile without spills) and the values are read in their stored order
// (Two more were built and measured on a B200 in round 2 and dropped -- profiles/r2a_variants.txt: values staged in
// shared memory with 16-byte cp.async, 23.8 vs 20.7 us per HVP; separate 16-bit natural positions for gather lengths
// beyond 2^20, 283 vs 288 us per 4-step recurrence at the news20 shape.)
enum GatherVariant { kGatherStored = 0, kGatherOrdered = 1, kGatherAsync = 2 };
// shared buffer of the per-segment sums; with asynchronous gathers the product buffer is busy receiving them, so
// the pieces of the deferred sums are staged here instead
constexpr int kSumCapAsync = kTileSegs + 256;

// ----------------------------------------------------------------------------- tile lists (host)
void free_segmat(SegMat *A) {
    cudaFree(A->ptr); cudaFree(A->idx); cudaFree(A->val); cudaFree(A->tiles); cudaFree(A->longtab);
    cudaFree(A->partial); cudaFree(A->longdot); cudaFree(A->ticket); cudaFree(A->segid); cudaFree(A->pos);
    cudaFree(A->gidx); cudaFree(A->gval);
    *A = SegMat();
}

// Tile list for one candidate tile size (tile_nnz non-zeros per CTA tile)
static void make_tiles(const std::vector<int32_t> &ptr, int64_t nseg, int64_t nnz, int tile_nnz, int short_seg,
                       std::vector<int4> &tiles, std::vector<int4> &longtab, int &npartial) {
    tiles.clear(); longtab.clear();
    tiles.reserve((size_t)(nnz / tile_nnz * 2 + nseg / kTileSegs + 16));
    npartial = 0;
    int64_t sg = 0;
    while (sg < nseg) {
        const int len = ptr[sg + 1] - ptr[sg];
        if (len > tile_nnz) {                                   // long segment: pieces + ordered combine
            const int pieces = (len + tile_nnz - 1) / tile_nnz;
            const int lid = (int)longtab.size();
            longtab.push_back(make_int4(npartial, pieces, ptr[sg], (int)sg));
            for (int p = 0; p < pieces; ++p) {
                const int k0 = ptr[sg] + p * tile_nnz;
                const int nk = std::min(tile_nnz, ptr[sg + 1] - k0);
                tiles.push_back(make_int4((int)sg, k0, -(lid + 1), nk));
            }
            npartial += pieces;
            ++sg;
            continue;
        }
        if (len > short_seg) {                                  // one medium segment: CTA-wide reduction
            tiles.push_back(make_int4((int)sg, ptr[sg], 1, len | (kSingleTag << 16)));
            ++sg;
            continue;
        }
        const int64_t s0 = sg;
        const int k0 = ptr[sg];
        int nk = 0, maxlen = 0;
        while (sg < nseg && sg - s0 < kTileSegs) {
            const int l = ptr[sg + 1] - ptr[sg];
            if (l > short_seg || nk + l > tile_nnz) break;
            nk += l;
            maxlen = std::max(maxlen, l);
            ++sg;
        }
        const int ns = (int)(sg - s0);
        int best = 0;
        long best_cost = -1;
        for (int lg = 0; lg <= 5; ++lg) {
            const int G = 1 << lg, groups = kSweepThreads >> lg;
            const long rounds = (ns + groups - 1) / groups;
            const long cost = rounds * ((maxlen + G - 1) / G + 2 * lg + 3);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = lg; }
        }
        tiles.push_back(make_int4((int)s0, k0, ns, nk | (best << 16)));
    }
}

// Gather order (setup, once per matrix): one CTA per tile sorts the entries of a stream tile by the index they
// gather with -- bitonic network over `index << kSlotBits | natural position` in shared memory; the words are
// unique, so the order is total -- and writes the words and the values in that order.  Single-segment tiles and
// pieces keep their natural order (sorted already for canonical input; their CTA-wide sums depend on it).
constexpr int kSortCap = 1 << kSlotBits;
__global__ void __launch_bounds__(kSweepThreads) gather_order_kernel(const int4 *__restrict__ tiles,
                                                                     const int32_t *__restrict__ idx,
                                                                     const double *__restrict__ val,
                                                                     uint32_t *__restrict__ gidx,
                                                                     double *__restrict__ gval) {
    __shared__ uint32_t key[kSortCap];
    const int4 td = tiles[blockIdx.x];
    const int k0 = td.y, nk = td.w & 0xffff, tag = td.w >> 16, tid = threadIdx.x;
    const bool stream = td.z > 0 && tag != kSingleTag;
    for (int i = tid; i < kSortCap; i += kSweepThreads)
        key[i] = i < nk ? ((uint32_t)idx[k0 + i] << kSlotBits) | (uint32_t)i : 0xffffffffu;
    __syncthreads();
    if (stream) {
        int np = 2;
        while (np < nk) np <<= 1;
        for (int k = 2; k <= np; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = tid; t < (np >> 1); t += kSweepThreads) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                    const uint32_t a = key[i], b = key[l];
                    if ((a > b) == ((i & k) == 0)) { key[i] = b; key[l] = a; }
                }
                __syncthreads();
            }
    }
    for (int i = tid; i < nk; i += kSweepThreads) {
        const uint32_t w = key[i];
        gidx[k0 + i] = w;
        gval[k0 + i] = val[k0 + (int)(w & (kSortCap - 1))];
    }
}

// The tile size is picked per matrix.  A sweep over a small (L2-resident) matrix is a handful of waves of CTA
// tiles, each wave one latency chain long: what counts is the NUMBER OF WAVES, e.g. the rcv1-shape CSR copy is 732
// tiles of 2048 -- two waves on 148 SMs x 4 CTAs, the second one a quarter full -- but 533 tiles of 2816: one wave
// (measured: 28.4 -> 25.2 us per HVP inside the Lanczos graph).  Among the candidates (8, 11, 12 non-zeros per
// thread) take the fewest waves, then the smaller tile; matrices of many waves keep 2048.
// Pure host logic (no device): the tile list of one matrix from its pointer array.  Also behind kcrn_plan_tiles, so
// that the CPU test suite can check its invariants on arbitrary segment-length distributions.
void plan_tiles(const std::vector<int32_t> &ptr, int64_t nseg, int64_t nnz, std::vector<int4> &tiles,
                std::vector<int4> &longtab, int &npartial, int &items, const Switches &sw) {
    const bool env_short = sw.short_seg > 0;                // tuning knob: longest segment that shares a tile
    const bool env_items = sw.tile_items > 0;               // tuning knob: force 8, 11 or 12 non-zeros per thread
    npartial = 0;
    // (three CTAs of 14 items per SM at 80 registers were measured as well: the row pass gains 0.7 us, the column copy no
    // longer fits one wave, 21.1 vs 20.3 us per HVP -- profiles/r2x_variants_three_ctas.txt)
    const int cand[3] = {8, 11, 12};
    const int slots = kNumSm * 4 - 32;                      // resident CTAs, minus a margin for the neighbours under PDL
    int short_seg = kShortSeg;
    items = 8;
    long best_waves = -1;
    for (int c : cand) {
        if (env_items && sw.tile_items != c) continue;
        const int tile_nnz = c * kSweepThreads;
        // few waves: also let segments of up to 2048 share tiles (fewer, fuller tiles; measured 24.6 -> 24.0 us per HVP)
        const int ss = env_short ? std::max(1, std::min(tile_nnz, sw.short_seg)) : 2 * kShortSeg;
        make_tiles(ptr, nseg, nnz, tile_nnz, c == 8 && !env_short ? kShortSeg : ss, tiles, longtab, npartial);
        long waves = ((long)tiles.size() + slots - 1) / slots;
        if (c == 8 && waves > 4 && !env_items) { items = 8; short_seg = env_short ? ss : kShortSeg; best_waves = waves; break; }
        if (c == 8 && !env_short) {              // the small-matrix packing for the default tile size as well
            make_tiles(ptr, nseg, nnz, tile_nnz, ss, tiles, longtab, npartial);
            waves = ((long)tiles.size() + slots - 1) / slots;
        }
        if (best_waves < 0 || waves < best_waves) { best_waves = waves; items = c; short_seg = ss; }
    }
    make_tiles(ptr, nseg, nnz, items * kSweepThreads, short_seg, tiles, longtab, npartial);
}

int build_tiles(SegMat *A, int64_t gather_len, cudaStream_t s, const Switches &sw) {
    A->gather_len = gather_len;
    const int64_t nseg = A->nseg;
    std::vector<int32_t> ptr((size_t)nseg + 1);
    KCRN_CUDA(cudaMemcpyAsync(ptr.data(), A->ptr, sizeof(int32_t) * (nseg + 1), cudaMemcpyDeviceToHost, s));
    KCRN_CUDA(cudaStreamSynchronize(s));
    std::vector<int4> tiles, longtab;
    int npartial = 0, items = 8;
    plan_tiles(ptr, nseg, A->nnz, tiles, longtab, npartial, items, sw);
    const int tile_nnz = items * kSweepThreads;
    A->items = items;
    if (sw.verbose)
        fprintf(stderr, "kcrn: %lld segments, %lld non-zeros -> %zu tiles of %d (%zu long segments)\n", (long long)nseg,
                (long long)A->nnz, tiles.size(), tile_nnz, longtab.size());
    A->ntiles = (int)tiles.size();
    A->nlong = (int)longtab.size();
    A->npartial = npartial;
    A->grid = std::max(1, A->ntiles);
    KCRN_CUDA(cudaMalloc(&A->tiles, sizeof(int4) * std::max<size_t>(tiles.size(), 1)));
    KCRN_CUDA(cudaMalloc(&A->longtab, sizeof(int4) * std::max<size_t>(longtab.size(), 1)));
    KCRN_CUDA(cudaMalloc(&A->partial, sizeof(double) * std::max(npartial, 1)));
    KCRN_CUDA(cudaMalloc(&A->longdot, sizeof(double) * (std::max(A->nlong, 1) + 1)));      // even length: 16-byte copies
    KCRN_CUDA(cudaMalloc(&A->ticket, sizeof(unsigned) * std::max(A->nlong, 1)));
    KCRN_CUDA(cudaMemsetAsync(A->ticket, 0, sizeof(unsigned) * std::max(A->nlong, 1), s));
    KCRN_CUDA(cudaMemsetAsync(A->longdot, 0, sizeof(double) * (std::max(A->nlong, 1) + 1), s));
    if (!tiles.empty())
        KCRN_CUDA(cudaMemcpyAsync(A->tiles, tiles.data(), sizeof(int4) * tiles.size(), cudaMemcpyHostToDevice, s));
    if (!longtab.empty())
        KCRN_CUDA(cudaMemcpyAsync(A->longtab, longtab.data(), sizeof(int4) * longtab.size(), cudaMemcpyHostToDevice, s));
    // gather-ordered copy for the sweeps (KCRN_NO_GATHER_ORDER=1: A/B switch, sweeps read idx / val as stored)
    if (A->nnz > 0 && A->ntiles > 0 && gather_len <= ((int64_t)1 << (32 - kSlotBits)) && tile_nnz <= kSortCap &&
        !sw.no_gather_order) {
        KCRN_CUDA(cudaMalloc(&A->gidx, sizeof(uint32_t) * (size_t)A->nnz));
        KCRN_CUDA(cudaMalloc(&A->gval, sizeof(double) * ((size_t)A->nnz + 2)));      // + 2: 16-byte pairs may straddle the end
        KCRN_CUDA(cudaMemsetAsync(A->gval + A->nnz, 0, sizeof(double) * 2, s));
        gather_order_kernel<<<(unsigned)A->ntiles, kSweepThreads, 0, s>>>(A->tiles, A->idx, A->val, A->gidx, A->gval);
        KCRN_LAUNCH_CHECK();
    }
    KCRN_CUDA(cudaStreamSynchronize(s));
    return KCRN_OK;
}

// ----------------------------------------------------------------------------- the sweep kernel
namespace {

enum SweepMode {
    kRowsForward = 0,   // A x, sigma, weights, residual, loss terms      (loss.py:215-232, 266-277, 296-297)
    kRowsWeighted = 1,  // t = D .* (A v)                                 (loss.py:299-301)
    kRowsLanczos = 2,   // same; v = source / beta from the control block, V[col] on the side, alpha = v'Hv (cubic.py:94,109)
    kColsRaw = 3,       // out = A' src                                   (row-sharded runs, before the allreduce)
    kColsScaled = 4,    // out = A' src / n (+ l2 u)                      (loss.py:227-229, :302)
    kColsLanczos = 5,   // w' = A' t / n + l2 v - beta v_prev - alpha v ; beta = |w'|   (loss.py:302 + cubic.py:93-97)
};

struct SweepParams {
    const int32_t *ptr, *idx;
    const double *val;
    const uint32_t *gidx;          // gather-ordered words and values of the tiles (null: idx / val as stored)
    const double *gval;
    const int4 *tiles;
    int ntiles;
    const int4 *longtab;
    const int32_t *segid;          // stored segment -> index of its operands / outputs (null: identity)
    double *partial, *longdot;
    unsigned *ticket;
    int nlong;
    const double *src;
    double *out;
    // forward
    const double *b;
    double *Ax, *D, *resid, *ell;
    // scaled columns
    const double *u;
    double n_samples, inv_n, l2;   // inv_n = 1 / n_samples (IEEE), for the three-operation exact quotient
    int add_l2;
    // Lanczos
    LanczosCtl *ctl;               // control block this kernel leaves behind (and reads, unless ctl_in differs)
    const LanczosCtl *ctl_in;      // state at the start of the kernel (deferred mode: the other half of the pair)
    int defer;                     // 1: grid-wide sums are finished by the CONSUMER kernel's prologue, not by a tail here
    int advance_here;              // rows-Lanczos, deferred: finish beta of the previous column sweep and advance
    const double *prev_part;       // per-CTA pieces left by the producer kernel (+ its long-segment pieces)
    const double *prev_longdot;
    int prev_count, prev_nlong;
    double *V;
    const double *g, *wp;
    double *bet;
    double *w, *alphas;
    int64_t d;
    double *red_partials;          // [2 * ntiles] per-tile contributions to the reduction of the pass
    unsigned *red_ticket;
    int alpha_mode;                // rows-Lanczos: 0 = finish alpha here, 1 = export the raw sums (row-sharded runs)
    double *alpha_raw, *alpha_raw2; // raw sums sum_i t_i (Av)_i (allreduced later) and |v|^2 for alpha_mode 1
    int update_ctl;                // cols-Lanczos: 1 = finish beta and advance the recurrence here
    // fused single-GPU loop (defer = 2) only:
#ifdef KCRN_SWEEP_TRACE
    unsigned long long *trace;     // [slots][1024 CTAs][12 marks]
    int trace_slot;
#endif
    int init_here, m_init;         // rows-Lanczos, first pass: |g| is still in the pieces the gradient sweep left; the
                                   // recurrence state of an m_init-step run starts here (no separate init kernel)
};

struct SweepState {
    const double *src = nullptr;
    double beta = 1.0, rinv = 1.0;        // gathered values are divided by beta (rows-Lanczos)
    const double *vcol = nullptr, *vpre = nullptr;
    double beta_prev = 0.0, alpha = 0.0;
    bool subtract = false;
};

constexpr int kSegRounds = kTileSegs / kSweepThreads;  // segments per thread in the epilogue
__device__ __forceinline__ int gather_index(int word, bool ordered) {
    return ordered ? (int)((unsigned)word >> kSlotBits) : word;
}
__host__ __device__ constexpr int sweep_operands(int mode) { return mode == kColsRaw ? 0 : (mode == kColsLanczos ? 2 : 1); }

// Epilogue of one finished segment.  op0 / op1 are the pass's per-segment operands, already in registers:
//   rows forward: label b            rows weighted: Hessian weight D
//   cols scaled : u (l2 term)        cols Lanczos : v[col], v[col-1]
template <int MODE>
__device__ __forceinline__ void finish_segment(const SweepParams &P, const SweepState &S, int seg, double sum,
                                               int lseg, double op0, double op1, double &acc) {
    if constexpr (MODE == kRowsForward) {
        const double a = sigmoid(sum);
        P.Ax[seg] = sum;
        P.D[seg] = __dmul_rn(a, 1.0 - a);
        P.resid[seg] = a - op0;
        const double el = __dsub_rn(__dmul_rn(1.0 - op0, sum), log_sigmoid(sum));
        P.ell[seg] = el;
        // share of the value (loss.py:215-220): summed per tile here, tiles in tile order by value_record_kernel
        if (lseg >= 0) P.longdot[lseg] = el;
        else acc += el;
    } else if constexpr (MODE == kRowsWeighted) {
        P.out[seg] = __dmul_rn(op0, sum);
    } else if constexpr (MODE == kRowsLanczos) {
        // t = D (Av); its share of alpha = v'Hv = sum_i D_i (Av)_i^2 / n (+ l2 |v|^2) is taken here, one sweep early
        const double t = __dmul_rn(op0, sum);
        P.out[seg] = t;
        const double prod = __dmul_rn(t, sum);
        if (lseg >= 0) P.longdot[lseg] = prod;  // long segments finish on whichever CTA arrives last:
        else acc += prod;                       // keep their share out of the per-tile sums
    } else if constexpr (MODE == kColsRaw) {
        P.out[seg] = sum;
    } else if constexpr (MODE == kColsScaled) {
        double y = quotient(sum, P.n_samples, P.inv_n);
        if (P.add_l2) y = __dadd_rn(y, __dmul_rn(P.l2, op0));
        P.out[seg] = y;
        if (P.red_partials) {                   // the gradient in front of a recurrence: pieces of |g|^2
            const double prod = __dmul_rn(y, y);
            if (lseg >= 0) P.longdot[lseg] = prod;
            else acc += prod;
        }
    } else {  // kColsLanczos
        double hv = __dadd_rn(quotient(sum, P.n_samples, P.inv_n), __dmul_rn(P.l2, op0));
        if (S.subtract) hv = __dsub_rn(hv, __dmul_rn(S.beta_prev, op1));
        hv = __dsub_rn(hv, __dmul_rn(S.alpha, op0));          // w' = Hv - beta v_prev - alpha v   (cubic.py:93-96)
        P.w[seg] = hv;
        const double prod = __dmul_rn(hv, hv);                // share of beta^2 = |w'|^2            (cubic.py:97)
        if (lseg >= 0) P.longdot[lseg] = prod;
        else acc += prod;
    }
}

template <int MODE>
__device__ __forceinline__ void load_operands(const SweepParams &P, const SweepState &S, int seg, double &op0,
                                              double &op1) {
    op0 = 0.0;
    op1 = 0.0;
    if constexpr (MODE == kRowsForward) op0 = __ldg(P.b + seg);
    if constexpr (MODE == kRowsWeighted || MODE == kRowsLanczos) op0 = __ldg(P.D + seg);
    if constexpr (MODE == kColsScaled) { if (P.add_l2) op0 = __ldg(P.u + seg); }
    if constexpr (MODE == kColsLanczos) {
        op0 = __ldg(S.vcol + seg);
        if (S.subtract) op1 = __ldg(S.vpre + seg);
    }
}

// fixed-order CTA sum: warp butterflies, then the warps in index order (thread 0 returns the total)
__device__ __forceinline__ double cta_sum_to_thread0(double v, double *sred) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) sred[warp] = v;
    __syncthreads();
    double tot = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kSweepWarps; ++w) tot += sred[w];
    return tot;
}

// two sums at once (same fixed order), one barrier
__device__ __forceinline__ void cta_sum2_to_thread0(double a, double b, double *sred, double &ra, double &rb) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) { sred[warp] = a; sred[kSweepWarps + warp] = b; }
    __syncthreads();
    ra = rb = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kSweepWarps; ++w) { ra += sred[w]; rb += sred[kSweepWarps + w]; }
}

// Deferred grid-wide sums: the producer kernel leaves one piece per CTA (and per long segment) and simply ends --
// no fence, no ticket, no last-CTA stage in its tail, which every CTA of every wave would otherwise pay for.  Each
// CTA of the consumer adds the pieces itself, in the fixed order of the old last-CTA stage (thread-strided, warp
// tree, warps in order), so all CTAs -- and all runs -- get the same bits.  The pieces travel global -> shared with
// cp.async (no registers, L2 only), issued as early as the data exists and collected as late as the result is
// needed, so that their latency hides behind the tile's own gathers.
//   a[n0] (and b[n0] right behind it at the even stride) + c[n1]; arrays are allocated to an even length.
__device__ __forceinline__ void defer_begin(const double *a, bool two, int n0, const double *c, int n1, double *sbuf) {
    const int n0e = (n0 + 1) & ~1, n1e = (n1 + 1) & ~1;
    for (int i = 2 * threadIdx.x; i < n0e; i += 2 * kSweepThreads) {
        cp_async_16(sbuf + i, a + i);
        if (two) cp_async_16(sbuf + n0e + i, a + n0e + i);
    }
    double *sc = sbuf + (two ? 2 : 1) * n0e;
    for (int i = 2 * threadIdx.x; i < n1e; i += 2 * kSweepThreads) cp_async_16(sc + i, c + i);
}

template <bool kAsyncGather>
__device__ __forceinline__ void defer_end(bool two, int n0, int n1, const double *sbuf, double *sred, double &ra,
                                          double &rb) {
    const int n0e = (n0 + 1) & ~1;
    if constexpr (kAsyncGather) cp_async_wait_group_1();      // the pieces; the gathers (last group) stay in flight
    else cp_async_wait_all();
    __syncthreads();
    double x = 0.0, y = 0.0;
    for (int i = threadIdx.x; i < n0; i += kSweepThreads) {
        x += sbuf[i];
        if (two) y += sbuf[n0e + i];
    }
    const double *sc = sbuf + (two ? 2 : 1) * n0e;
    for (int i = threadIdx.x; i < n1; i += kSweepThreads) x += sc[i];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    x = warp_sum(x);
    y = warp_sum(y);
    if (lane == 0) { sred[warp] = x; sred[kSweepWarps + warp] = y; }
    __syncthreads();
    ra = rb = 0.0;
#pragma unroll
    for (int w = 0; w < kSweepWarps; ++w) { ra += sred[w]; rb += sred[kSweepWarps + w]; }
    __syncthreads();      // sbuf (the product buffer of the tile) and sred are free again
}

template <int MODE, int kItems, int GATHER>
__global__ void __launch_bounds__(kSweepThreads, 4) sweep_kernel(const SweepParams P) {
    constexpr int kTileNnz = kItems * kSweepThreads;       // non-zeros per CTA tile, kItems per thread
    constexpr bool ordered = GATHER != kGatherStored, kAsyncGather = GATHER == kGatherAsync;
    __shared__ __align__(16) double sprod[kTileNnz];       // products of the tile (16-byte aligned: the deferred sums' pieces arrive by 16-byte cp.async)
    __shared__ __align__(16) double ssum[kAsyncGather ? kSumCapAsync : kTileSegs];   // per-segment sums of the tile
    double *const spieces = kAsyncGather ? ssum : sprod;   // staging of the deferred sums' pieces (the product buffer is
                                                           // busy receiving the asynchronous gathers)
    __shared__ int sptr[kTileSegs + 8];          // segment pointers of the tile
    __shared__ double sred[2 * kSweepWarps];
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31;
    constexpr int NOP = sweep_operands(MODE);

    // Everything that only depends on the (constant) matrix is requested BEFORE waiting for the previous kernel
    // of the stream: with programmatic dependent launch this CTA may already be resident while that kernel is
    // still finishing, so the descriptor -> (index, value) part of the latency chain overlaps with its tail.
    const int tile = blockIdx.x;
    const unsigned nblk = gridDim.x;
    KTRACE(0);                       // kernel entry
    int4 td = make_int4(0, 0, 0, 0);
    int ci[kItems];                  // index, or gather-ordered word (index << kSlotBits | natural position); -1: none
    double av[kItems];
    if (tile < P.ntiles) {
        // (The tile's first entry and length were also tried in the kernel's parameter bank, so that the stream is
        // requested without waiting for this descriptor load: no gain, profiles/r2v_variants_param_desc.txt.)
        td = __ldg(P.tiles + tile);
        const int k0s = td.y, nk0 = td.w & 0xffff;
        const int32_t *ip = (ordered ? reinterpret_cast<const int32_t *>(P.gidx) : P.idx) + k0s;
        const double *vp = (ordered && !kAsyncGather ? P.gval : P.val) + k0s;
        // all index words first, then the values: the gathers only need the words (a third of the stream's bytes), so
        // they can start while the values are still on their way (-DKCRN_INTERLEAVED_STREAM: the old order, for A/B)
#ifdef KCRN_INTERLEAVED_STREAM
#pragma unroll
        for (int r = 0; r < kItems; ++r) {
            const int i = r * kSweepThreads + tid;
            ci[r] = i < nk0 ? ld_stream(ip + i) : -1;
            av[r] = i < nk0 ? ld_stream(vp + i) : 0.0;
        }
#else
#pragma unroll
        for (int r = 0; r < kItems; ++r) {
            const int i = r * kSweepThreads + tid;
            ci[r] = i < nk0 ? ld_stream(ip + i) : -1;
        }
#pragma unroll
        for (int r = 0; r < kItems; ++r) {
            const int i = r * kSweepThreads + tid;
            av[r] = i < nk0 ? ld_stream(vp + i) : 0.0;
        }
#endif
    }
    if (tile < P.ntiles) {
        if (td.z > 0 && (td.w >> 16) != kSingleTag) {
#pragma unroll
            for (int r = 0; r < kSegRounds + 1; ++r) {
                const int i = r * kSweepThreads + tid;
                if (i <= td.z) cp_async_4(sptr + i, P.ptr + td.x + i);
            }
        }
    }
    KTRACE(1);                       // matrix-only loads requested
    pdl_wait();      // the previous kernel's results (control block, vectors) are complete and visible from here on
    KTRACE(2);                       // predecessor complete

    SweepState S;
    S.src = P.src;
    bool final_ = false;
    int col = 0, store_idx = 0, redo = 0;
    double xv[kItems];                   // gathered values of the tile
    bool gathered = false;
    // the gathers of the tile from `src`: into registers, or (asynchronous) straight into the natural slots of sprod
    auto gather_from = [&](const double *src) {
#pragma unroll
        for (int r = 0; r < kItems; ++r) {
            if constexpr (kAsyncGather) {
                const int i = r * kSweepThreads + tid;
                if (ci[r] != -1)
                    cp_async_8(sprod + KCRN_SLOT(ordered ? (ci[r] & ((1 << kSlotBits) - 1)) : i, i), src + gather_index(ci[r], ordered));
            } else {
                xv[r] = ci[r] != -1 ? __ldg(src + gather_index(ci[r], ordered)) : 0.0;
            }
        }
        if constexpr (kAsyncGather) cp_async_commit();
    };
    // before a CTA leaves early nothing may still be on its way into its shared memory
    auto drain = [&]() { cp_async_wait_all(); };
    bool alpha_pending = false;
    LanczosCtl st_cols;
    if constexpr (MODE == kRowsLanczos) {
        const bool scribe = tile == 0 && tid == 0;            // the one thread that persists what all compute
        // Requests that do not depend on the control block go first: the pieces of the sum this kernel has to finish
        // (beta = |w'| of the column sweep before us, or |g| of the gradient sweep).  (Starting the gathers here as
        // well, speculatively from the vector the host knows, was measured: no gain, profiles/r2g_ab.txt.)
        const bool pieces = P.advance_here || P.init_here;
        if (pieces) {
            defer_begin(P.prev_part, false, P.prev_count, P.prev_longdot, P.prev_nlong, spieces);
            if constexpr (kAsyncGather) cp_async_commit();
        }
        LanczosCtl st;
        if (P.init_here) {
            // a fresh recurrence (what lanczos_init_kernel does): bet[0] = |g|, everything else reset
            if (!gathered) { gather_from(P.g); gathered = true; }
            double gsq, unused;
            defer_end<kAsyncGather>(false, P.prev_count, P.prev_nlong, spieces, sred, gsq, unused);
            const double gn = sqrt(gsq);
            st.col = 0; st.redo = 0; st.final_ = (P.m_init == 1); st.done = 0; st.m_eff = P.m_init; st.nvalid = 1;
            st.store_idx = P.m_init - 1; st.m = P.m_init; st.broke = 0; st.hvps = 0; st.last_beta = 0.0;
            st.cur_beta = gn; st.cur_alpha = 0.0;
            if (tile == 0) {
                for (int i = tid; i < P.m_init; i += kSweepThreads) { P.alphas[i] = 0.0; P.bet[i + 1] = 0.0; }
                if (tid == 0) P.bet[0] = gn;
            }
        } else {
            st = *P.ctl_in;
            if (st.done) {
                drain();                                       // nothing may still be on its way into this CTA's shared memory
                if (scribe && P.ctl != P.ctl_in) *P.ctl = st;      // keep the pair in step
                return;
            }
            if (P.advance_here) {
                // beta = |w'| of the column sweep before us is still in pieces.  While they travel, gather from w'
                // -- the source unless the recurrence breaks down right now, which the sum will tell
                KTRACE(10);                                    // control block has arrived
                if (!gathered) { gather_from(P.wp); gathered = true; }
                KTRACE(11);                                    // index words have arrived, gathers are out
                double bsq, unused;
                defer_end<kAsyncGather>(false, P.prev_count, P.prev_nlong, spieces, sred, bsq, unused);
                lanczos_advance(&st, scribe ? P.bet : nullptr, sqrt(bsq));     // cubic.py:98-109, in registers
            }
            if (st.redo && gathered) {                         // breakdown: the operator is re-applied to V[col]
                gathered = false;
                if constexpr (kAsyncGather) { cp_async_wait_all(); __syncthreads(); }   // the slots are written again
            }
        }
        if (scribe && P.ctl != P.ctl_in) *P.ctl = st;
        col = st.col;
        redo = st.redo;
        final_ = st.final_ != 0;
        store_idx = st.store_idx;
        if (redo) {
            S.src = P.V + (size_t)col * P.d;
        } else {
            S.src = col == 0 ? P.g : P.wp;
            S.beta = st.cur_beta;                // = bet[col], without a dependent second load
            S.rinv = 1.0 / S.beta;
        }
    }
    if constexpr (MODE == kColsLanczos) {
        // first what does not depend on the control block: the pieces of alpha = v'Hv the row sweep before us left (needed
        // by the epilogue only)
        if (P.defer) {
            defer_begin(P.prev_part, true, P.prev_count, P.prev_longdot, P.prev_nlong, spieces);
            if constexpr (kAsyncGather) cp_async_commit();
            alpha_pending = true;
        }
        st_cols = *P.ctl_in;
        if (st_cols.done) { drain(); return; }
        KTRACE(10);                                            // control block has arrived
        col = st_cols.col;
        if (P.defer) {
            if (st_cols.final_) {                              // cubic.py:109: the last alpha needs no further vector
                double s1, s2;
                defer_end<false>(true, P.prev_count, P.prev_nlong, spieces, sred, s1, s2);      // waits for everything in flight
                drain();
                if (tile == 0 && tid == 0) {
                    const double alpha = __dadd_rn(quotient(s1, P.n_samples, P.inv_n), __dmul_rn(P.l2, s2));
                    P.alphas[st_cols.store_idx] = alpha;
                    P.ctl->cur_alpha = alpha;
                    P.ctl->hvps = st_cols.hvps + 1;
                    P.ctl->done = 1;
                }
                return;
            }
        } else {
            S.alpha = st_cols.cur_alpha;         // = alphas[col], written by the row sweep of this step
        }
        S.vcol = P.V + (size_t)col * P.d;
        if (col > 0) {
            S.subtract = true;
            S.vpre = P.V + (size_t)(col - 1) * P.d;
            S.beta_prev = st_cols.cur_beta;      // = bet[col]
        }
    }

    KTRACE(3);                       // control block known (and the sum this kernel had to finish)
    double acc = 0.0;
    double vsq = 0.0;
    if constexpr (MODE == kRowsLanczos) {
        // side job, issued before the tile so that its loads overlap the gathers: store the normalised vector as
        // basis column `col` (the sweep itself divides on the fly); |v|^2 for the l2 part of alpha
        if (!redo || P.l2 != 0.0) {
            double *vout = P.V + (size_t)col * P.d;
            for (int64_t c = (int64_t)tile * blockDim.x + tid; c < P.d; c += (int64_t)nblk * blockDim.x) {
                const double vc = quotient(__ldg(S.src + c), S.beta, S.rinv);
                if (!redo) vout[c] = vc;
                vsq = fma(vc, vc, vsq);
            }
        }
    }
    if (tile < P.ntiles) {
        const int seg0 = td.x, k0 = td.y, z = td.z, nk = td.w & 0xffff, tag = td.w >> 16;
        const bool stream = z > 0 && tag != kSingleTag;
        int4 lt = make_int4(0, 1, 0, 0);
        // epilogue operands of the FIRST segment round to registers: in flight while the gathers run.  (Tiles of more
        // than 256 segments -- runs of tiny columns -- fetch the operands of their later rounds when they need them:
        // four rounds of prefetched operands cost the Lanczos passes 9 to 15 registers they do not have.)
        double o0 = 0.0, o1 = 0.0;
        double e0 = 0.0, e1 = 0.0;           // operands of a single-segment tile (thread 0)
        int sid0 = seg0;
        if (stream) {
            if (tid < z) {
                sid0 = P.segid ? __ldg(P.segid + seg0 + tid) : seg0 + tid;
                load_operands<MODE>(P, S, sid0, o0, o1);
            }
        } else {
            if (z < 0) lt = __ldg(P.longtab + (-z - 1));
            if (tid == 0) {
                if (P.segid) sid0 = __ldg(P.segid + seg0);
                load_operands<MODE>(P, S, sid0, e0, e1);
            }
        }
        KTRACE(4);                   // basis column stored, operands requested
        if (!gathered) gather_from(S.src);
        if constexpr (MODE == kColsLanczos) {
            if (alpha_pending) {                 // the gathers above are in flight while the pieces are added
                double s1, s2;
                defer_end<kAsyncGather>(true, P.prev_count, P.prev_nlong, spieces, sred, s1, s2);
                S.alpha = __dadd_rn(quotient(s1, P.n_samples, P.inv_n), __dmul_rn(P.l2, s2));
                if (tile == 0 && tid == 0) {
                    P.alphas[col] = S.alpha;
                    P.ctl->cur_alpha = S.alpha;
                    P.ctl->hvps = st_cols.hvps + 1;
                }
            }
        }
        if constexpr (kAsyncGather) {        // the gathered values are in their natural slots: collect this thread's
            cp_async_wait_all();
            __syncthreads();
#pragma unroll
            for (int r = 0; r < kItems; ++r) {
                const int i = r * kSweepThreads + tid;
                xv[r] = i < nk ? sprod[i] : 0.0;
            }
        }
        KTRACE(5);                   // gathered values have arrived (and alpha, columns)
        double p[kItems];
#pragma unroll
        for (int r = 0; r < kItems; ++r) {
            double x = xv[r];
            if constexpr (MODE == kRowsLanczos) x = quotient(x, S.beta, S.rinv);      // beta = 1 when re-applying
            p[r] = av[r] * x;
        }
        if (!stream) {
            // one segment (or one piece of a long one) spread over the whole CTA
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < kItems; ++r) s += p[r];
            const double tot = cta_sum_to_thread0(s, sred);
            if (z > 0) {
                if (tid == 0) finish_segment<MODE>(P, S, sid0, tot, -1, e0, e1, acc);
            } else {
                const int lseg = -z - 1;
                if (tid == 0) {
                    P.partial[lt.x + (k0 - lt.z) / kTileNnz] = tot;
                    __threadfence();
                    s_last = atomicInc(P.ticket + lseg, (unsigned)(lt.y - 1)) == (unsigned)(lt.y - 1);
                }
                __syncthreads();
                if (s_last && tid < 32) {
                    __threadfence();
                    double all = 0.0;
                    for (int i = lane; i < lt.y; i += 32) all += __ldcg(P.partial + lt.x + i);
                    all = warp_sum(all);
                    // thread 0 holds the operands
                    if (tid == 0) finish_segment<MODE>(P, S, sid0, all, lseg, e0, e1, acc);
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < kItems; ++r) {
                const int i = r * kSweepThreads + tid;
                if (i < nk) sprod[KCRN_SLOT(GATHER == kGatherOrdered ? (ci[r] & ((1 << kSlotBits) - 1)) : i, i)] = p[r];   // the natural position
            }
            cp_async_wait_all();
            __syncthreads();
            KTRACE(6);               // products in their natural slots
            // phase A: every group of G threads sums one segment from shared memory, left to right
            const int G = 1 << tag, groups = kSweepThreads >> tag, gid = tid >> tag, gl = tid & (G - 1);
            for (int sb = 0; sb < z; sb += groups) {
                const int sg = sb + gid;
                int a = 0, b = 0;
                if (sg < z) { a = sptr[sg] - k0; b = sptr[sg + 1] - k0; }
                double sum = 0.0;
                for (int i = a + gl; i < b; i += G) sum += sprod[i];
                for (int o = G >> 1; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                if (sg < z && gl == 0) ssum[sg] = sum;
            }
            __syncthreads();
            KTRACE(7);               // segment sums
            // phase B: one thread per segment applies the epilogue, coalesced
#pragma unroll
            for (int r = 0; r < kSegRounds; ++r) {
                const int sg = r * kSweepThreads + tid;
                if (sg < z) {
                    int sidr = sid0;
                    double a0 = o0, a1 = o1;
                    if (r > 0) {
                        sidr = P.segid ? __ldg(P.segid + seg0 + sg) : seg0 + sg;
                        load_operands<MODE>(P, S, sidr, a0, a1);
                    }
                    finish_segment<MODE>(P, S, sidr, ssum[sg], -1, a0, a1, acc);
                }
            }
        }
    }

    // The tile is done: from here on only the short reduction tails remain, so the next kernel of the stream may
    // become resident and start its own (matrix-only) loads while they run -- it still waits for our completion.
    KTRACE(8);                       // epilogue stored
    pdl_trigger();

    if constexpr (MODE == kColsScaled) {
        if (P.red_partials) {
            __syncthreads();
            const double mine = cta_sum_to_thread0(acc, sred);
            if (tid == 0) P.red_partials[tile] = mine;
        }
    }

    if constexpr (MODE == kRowsForward) {
        __syncthreads();
        const double mine = cta_sum_to_thread0(acc, sred);
        if (tid == 0) P.red_partials[tile] = mine;
    }

    if constexpr (MODE == kRowsLanczos) {
        // alpha = v'Hv: threads -> tile (fixed order) -> tiles in tile order, by the consumer (deferred) or by the
        // last CTA here
        __syncthreads();
        double mine, mine2;
        cta_sum2_to_thread0(acc, vsq, sred, mine, mine2);
        if (tid == 0) {
            P.red_partials[tile] = mine;
            P.red_partials[((nblk + 1) & ~1u) + tile] = mine2;            // even stride: 16-byte copies later
        }
        if (P.defer) { KTRACE(9); return; }
        if (tid == 0) {
            __threadfence();
            s_last = atomicInc(P.red_ticket, nblk - 1) == nblk - 1;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            double tot = 0.0, tot2 = 0.0;
            for (int i = tid; i < (int)nblk; i += kSweepThreads) {
                tot += __ldcg(P.red_partials + i);
                tot2 += __ldcg(P.red_partials + ((nblk + 1) & ~1u) + i);
            }
            for (int i = tid; i < P.nlong; i += kSweepThreads) tot += __ldcg(P.longdot + i);
            __syncthreads();
            double s1, s2;
            cta_sum2_to_thread0(tot, tot2, sred, s1, s2);
            if (tid == 0) {
                if (P.alpha_mode == 0) {
                    const double alpha = __dadd_rn(quotient(s1, P.n_samples, P.inv_n), __dmul_rn(P.l2, s2));
                    P.alphas[final_ ? store_idx : col] = alpha;
                    P.ctl->cur_alpha = alpha;
                    P.ctl->hvps += 1;
                    if (final_) P.ctl->done = 1;           // cubic.py:109: the last alpha needs no further vector
                } else {
                    P.alpha_raw[0] = s1;                   // summed over the ranks together with A't
                    P.alpha_raw2[0] = s2;
                }
            }
        }
    }

    if constexpr (MODE == kColsLanczos) {
        // beta^2 = |w'|^2: threads -> tile (fixed order) -> tiles in tile order, by the next row sweep (deferred) or
        // by the last CTA here
        if (!P.update_ctl) { return; }    // re-orthogonalisation follows: it recomputes the norm
        __syncthreads();
        const double mine = cta_sum_to_thread0(acc, sred);
        if (tid == 0) P.red_partials[tile] = mine;
        if (P.defer) { KTRACE(9); return; }
        if (tid == 0) {
            __threadfence();
            s_last = atomicInc(P.red_ticket, nblk - 1) == nblk - 1;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            double tot = 0.0;
            for (int i = tid; i < (int)nblk; i += kSweepThreads) tot += __ldcg(P.red_partials + i);
            for (int i = tid; i < P.nlong; i += kSweepThreads) tot += __ldcg(P.longdot + i);
            __syncthreads();
            const double bsq = cta_sum_to_thread0(tot, sred);
            if (tid == 0) lanczos_advance(P.ctl, P.bet, sqrt(bsq));
        }
    }
    KTRACE(9);                       // tail sums written
}

void fill_matrix(SweepParams &P, const SegMat &A) {
    P.ptr = A.ptr; P.idx = A.idx; P.val = A.val; P.gidx = A.gidx; P.gval = A.gval; P.tiles = A.tiles; P.ntiles = A.ntiles;
    P.longtab = A.longtab; P.segid = A.segid; P.partial = A.partial; P.longdot = A.longdot; P.ticket = A.ticket;
    P.nlong = A.nlong;
}

template <int MODE, int GATHER>
int launch_variant(kcrn_problem *h, const SegMat &A, const SweepParams &P, int grid) {
    if (A.items == 11) KCRN_CUDA(launch_kernel(sweep_kernel<MODE, 11, GATHER>, dim3(grid), dim3(kSweepThreads), 0, h->stream, h->pdl, P));
    else if (A.items == 12) KCRN_CUDA(launch_kernel(sweep_kernel<MODE, 12, GATHER>, dim3(grid), dim3(kSweepThreads), 0, h->stream, h->pdl, P));
    else KCRN_CUDA(launch_kernel(sweep_kernel<MODE, 8, GATHER>, dim3(grid), dim3(kSweepThreads), 0, h->stream, h->pdl, P));
    return KCRN_OK;
}

// The gather variant is a property of (matrix, class of pass): the Lanczos passes carry more state than the others
template <int MODE>
int launch(kcrn_problem *h, const SegMat &A, SweepParams &P, double *longdot = nullptr) {
    fill_matrix(P, A);
    if (longdot) P.longdot = longdot;      // the forward sweep keeps its long-segment shares apart from the Lanczos passes'
    const int variant = !A.gidx ? kGatherStored : (MODE == kRowsLanczos || MODE == kColsLanczos ? A.gather_lanczos : A.gather_plain);
    auto go = [&](int grid) -> int {
        if (variant == kGatherAsync) KCRN_TRY((launch_variant<MODE, kGatherAsync>(h, A, P, grid)));
        else if (variant == kGatherOrdered) KCRN_TRY((launch_variant<MODE, kGatherOrdered>(h, A, P, grid)));
        else KCRN_TRY((launch_variant<MODE, kGatherStored>(h, A, P, grid)));
        h->counters[0] += 1;
        return KCRN_OK;
    };
    return go(A.grid);
}

// everything that loss.py derives from a cached A x, for kcrn_set_Ax (SSCN's incremental update)
__global__ void derive_from_Ax_kernel(const double *__restrict__ Ax, const double *__restrict__ b, int64_t n,
                                      double *D, double *resid, double *ell) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double z = Ax[i], bi = b[i];
    const double a = sigmoid(z);
    D[i] = __dmul_rn(a, 1.0 - a);
    resid[i] = a - bi;
    ell[i] = __dsub_rn(__dmul_rn(1.0 - bi, z), log_sigmoid(z));
}

}  // namespace

int launch_forward(kcrn_problem *h, const double *x) {
    SweepParams P{};
    P.src = x; P.b = h->b; P.Ax = h->Ax; P.D = h->D; P.resid = h->resid; P.ell = h->ell;
    P.red_partials = h->fwd_part;
    h->counters[3] += 1;
    h->fwd_parts_valid = true;
    return launch<kRowsForward>(h, h->csr, P, h->fwd_longdot);
}

int launch_rows_weighted(kcrn_problem *h, const double *v) {
    SweepParams P{};
    P.src = v; P.D = h->D; P.out = h->t;
#ifdef KCRN_SWEEP_TRACE
    P.trace = h->sweep_trace; P.trace_slot = 20;      // the plain passes next to the Lanczos ones
#endif
    h->counters[3] += 1;
    return launch<kRowsWeighted>(h, h->csr, P);
}

int gather_candidates_last() { return kGatherAsync; }

// Which gather variant serves the passes outside the Lanczos recurrence: measured per matrix on the weighted row
// pass / the scaled column pass, a few back-to-back launches each (the operands are zeroed scratch; all variants
// compute the same bits, so the choice never shows in a result).  The Lanczos passes are measured at their first
// use (kcrn_api.cu: tune_lanczos_sweeps), when a gradient to start the recurrence from exists.
int tune_plain_sweeps(kcrn_problem *h) {
    const int fp = h->sw.gather_plain, fl = h->sw.gather_lanczos;      // KCRN_GATHER (A/B switch); -1: measure
    for (SegMat *A : {&h->csr, &h->csc}) {
        A->gather_plain = A->gather_lanczos = kGatherStored;
        if (!A->gidx) continue;
        A->gather_plain = fp >= 0 ? fp : kGatherOrdered;
        A->gather_lanczos = fl >= 0 ? fl : kGatherAsync;     // until measured (row-sharded runs never measure)
    }
    if (fp >= 0 || (!h->csr.gidx && !h->csc.gidx)) return KCRN_OK;
    cudaStream_t s = h->stream;
    KCRN_CUDA(cudaMemsetAsync(h->x_cur, 0, sizeof(double) * h->d, s));
    KCRN_CUDA(cudaMemsetAsync(h->D, 0, sizeof(double) * h->n, s));
    KCRN_CUDA(cudaMemsetAsync(h->t, 0, sizeof(double) * h->n, s));
    int64_t saved[8];
    memcpy(saved, h->counters, sizeof(saved));
    cudaEvent_t e0, e1;
    KCRN_CUDA(cudaEventCreate(&e0));
    KCRN_CUDA(cudaEventCreate(&e1));
    int rc = KCRN_OK;
    for (int which = 0; which < 2 && rc == KCRN_OK; ++which) {
        SegMat *A = which == 0 ? &h->csr : &h->csc;
        if (!A->gidx || A->nnz == 0) continue;
        const int reps = A->nnz > 50000000 ? 3 : 10;
        int best = kGatherStored;
        float best_ms = 0.f;
        const int batches = A->nnz > 50000000 ? 1 : 3;              // the fastest of a few batches: robust to a stray stall
        for (int v = kGatherStored; v <= gather_candidates_last() && rc == KCRN_OK; ++v) {
            A->gather_plain = v;
            float ms = 0.f;
            for (int bt = 0; bt < batches && rc == KCRN_OK; ++bt) {
                for (int i = (bt == 0 ? -2 : 0); i < reps && rc == KCRN_OK; ++i) {      // two warm-up launches first
                    if (i == 0) cudaEventRecord(e0, s);
                    rc = which == 0 ? launch_rows_weighted(h, h->x_cur) : launch_cols_scaled(h, h->t, nullptr, h->o_stage, nullptr);
                }
                cudaEventRecord(e1, s);
                if (cudaEventSynchronize(e1) != cudaSuccess) { set_error("tune_plain_sweeps: %s", cudaGetErrorString(cudaGetLastError())); rc = KCRN_E_CUDA; }
                float t = 0.f;
                cudaEventElapsedTime(&t, e0, e1);
                if (bt == 0 || t < ms) ms = t;
            }
            if (v == kGatherStored || ms < best_ms) { best = v; best_ms = ms; }
            if (h->sw.verbose)
                fprintf(stderr, "kcrn: %s pass, gather variant %d: %.2f us\n", which == 0 ? "row" : "column", v, 1e3 * ms / reps);
        }
        A->gather_plain = best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    memcpy(h->counters, saved, sizeof(saved));
    return rc;
}

// The consumer-side sums stage the producer's pieces in the tile's product buffer: both sets must fit
bool lanczos_defer_fits(const kcrn_problem *h) {
    auto even = [](int v) { return (v + 1) & ~1; };
    if (h->sharded()) {                    // only the row sweep stages pieces (those of the epilogue kernel); the alpha pieces
                                           // are added by the epilogue kernel straight from global memory
        const bool async_r = h->csr.gidx && h->csr.gather_lanczos >= kGatherAsync;
        return even(h->vec_grid) <= (async_r ? kSumCapAsync : h->csr.items * kSweepThreads);
    }
    // staged in the CONSUMER's product buffer: row pieces in the column sweep's, column pieces in the row sweep's
    const bool async_cols = h->csc.gidx && h->csc.gather_lanczos >= kGatherAsync;     // async / staged: pieces go to the sums buffer
    const bool async_rows = h->csr.gidx && h->csr.gather_lanczos >= kGatherAsync;
    const int cap_cols = async_cols ? kSumCapAsync : h->csc.items * kSweepThreads;
    const int cap_rows = async_rows ? kSumCapAsync : h->csr.items * kSweepThreads;
    return 2 * even(h->csr.grid) + even(h->csr.nlong) <= cap_cols && even(h->csc.grid) + even(h->csc.nlong) <= cap_rows;
}

// slot = position of the step in the captured recurrence; defer = 1 (single-GPU fused loop): grid-wide sums are
// finished by the consumer kernel and the control block ping-pongs between h->ctl[0] and h->ctl[1] so that the
// row sweep can advance the recurrence in registers while other CTAs still read the old state.
int launch_rows_lanczos(kcrn_problem *h, double *alpha_raw, int slot, int defer, int init_m) {
    SweepParams P{};
    P.D = h->D; P.out = h->t; P.V = h->V; P.g = h->g; P.wp = h->wp; P.bet = h->bet; P.d = h->d;
    P.alphas = h->alphas; P.n_samples = (double)h->n_global; P.inv_n = 1.0 / P.n_samples; P.l2 = h->l2;
    P.alpha_mode = h->sharded() ? 1 : 0; P.alpha_raw = alpha_raw; P.alpha_raw2 = h->scal + 17;   // the first is summed over the ranks with A't
    P.defer = defer;
    if (defer == 2) {                      // advance here: beta pieces of the previous column sweep
        P.ctl_in = h->ctl + (slot & 1); P.ctl = h->ctl + ((slot + 1) & 1);
        P.advance_here = slot > 0;
        if (h->sharded()) {                // row-sharded: the epilogue kernel after the exchange left the pieces of |w'|^2
            P.prev_part = h->post_part; P.prev_count = h->vec_grid; P.prev_longdot = h->post_part; P.prev_nlong = 0;
        } else {
            P.prev_part = h->lan_cols_part; P.prev_count = h->csc.grid; P.prev_longdot = h->csc.longdot; P.prev_nlong = h->csc.nlong;
        }
        if (slot == 0 && init_m > 0) { P.init_here = 1; P.m_init = init_m; }
    } else {
        P.ctl_in = P.ctl = h->ctl;
    }
    P.red_partials = defer ? h->lan_rows_part : h->red_partials; P.red_ticket = h->red_ticket + 7;
#ifdef KCRN_SWEEP_TRACE
    P.trace = h->sweep_trace; P.trace_slot = 2 * slot;
#endif
    h->counters[3] += 1;
    return launch<kRowsLanczos>(h, h->csr, P);
}

int launch_cols_raw(kcrn_problem *h, const double *src, double *out) {
    SweepParams P{};
    P.src = src; P.out = out;
    h->counters[4] += 1;
    return launch<kColsRaw>(h, h->csc, P);
}

int launch_cols_scaled(kcrn_problem *h, const double *src, const double *u, double *out, double *sq_part) {
    SweepParams P{};
    P.src = src; P.out = out; P.u = u; P.n_samples = (double)h->n_global; P.inv_n = 1.0 / P.n_samples; P.l2 = h->l2;
    P.add_l2 = u != nullptr;
    P.red_partials = sq_part;              // pieces of |out|^2 per tile (+ csc.longdot), for the recurrence that follows
#ifdef KCRN_SWEEP_TRACE
    P.trace = h->sweep_trace; P.trace_slot = 21;
#endif
    h->counters[4] += 1;
    return launch<kColsScaled>(h, h->csc, P);
}

int launch_cols_lanczos(kcrn_problem *h, int update_ctl, int slot, int defer) {
    SweepParams P{};
    P.src = h->t; P.V = h->V; P.bet = h->bet; P.w = h->wp; P.alphas = h->alphas; P.d = h->d;
    P.update_ctl = update_ctl;
    P.n_samples = (double)h->n_global; P.inv_n = 1.0 / P.n_samples; P.l2 = h->l2;
    P.defer = defer;
    P.ctl_in = P.ctl = defer == 2 ? h->ctl + ((slot + 1) & 1) : h->ctl;
    if (defer) {
        P.prev_part = h->lan_rows_part; P.prev_count = h->csr.grid; P.prev_longdot = h->csr.longdot; P.prev_nlong = h->csr.nlong;
    }
    P.red_partials = defer ? h->lan_cols_part : h->red_partials; P.red_ticket = h->red_ticket + 0;
#ifdef KCRN_SWEEP_TRACE
    P.trace = h->sweep_trace; P.trace_slot = 2 * slot + 1;
#endif
    h->counters[4] += 1;
    return launch<kColsLanczos>(h, h->csc, P);
}

int launch_derive_from_Ax(kcrn_problem *h) {
    if (h->n == 0) return KCRN_OK;
    derive_from_Ax_kernel<<<(unsigned)((h->n + 255) / 256), 256, 0, h->stream>>>(h->Ax, h->b, h->n, h->D,
                                                                                h->resid, h->ell);
    KCRN_LAUNCH_CHECK();
    h->counters[0] += 1;
    return KCRN_OK;
}

}  // namespace kcrn
