// The whole Lanczos recurrence as ONE persistent kernel (small problems: C1 / C2 of BASELINE.json).
//
// At the rcv1 shape a pass over A is 1.5 M non-zeros -- 10 k per SM -- and the graph of per-pass kernels spends more
// time on kernel boundaries, control-block round trips and re-staging than on the non-zeros (20.7 us per HVP against
// 5.8 us of HBM roofline, profiles/r2a_variants.txt).  This kernel keeps one 1024-thread CTA resident on every SM for
// the whole m-step recurrence (optimizer/cubic.py:77-111):
//
//   * every CTA owns ONE tile of the CSR copy and ONE tile of the CSC copy (balanced: nnz / SMs entries each, cut
//     segments finish through piece sums + a ticket, in piece order -- no floating-point atomics);
//   * a pass = wait for the tile's stream -> gather + multiply in place -> segment sums -> epilogue -> grid barrier.
//     The tile stream (32-bit gather-ordered words, 64-bit values) is constant, so the NEXT pass's tile is brought
//     into shared memory by bulk copies (cp.async.bulk + mbarrier, SASS UBLKCP) issued as soon as the buffers
//     of the current pass are consumed: the stream never waits behind a kernel launch and holds no registers;
//   * the words of a tile are sorted by gather index over ~10 k entries, so a warp-wide gather touches 3-4 distinct
//     128-byte lines instead of 12-16 (2816-entry tiles) or 31 (stored order): the L1TEX tag stage, which bounds the
//     per-pass kernels, sees a quarter of the requests;
//   * the d-vectors of the recurrence (w', the basis) live in the stored column order of the CSC copy, so the
//     column pass reads v, v_prev and writes w' coalesced, and the row pass gathers popular columns from adjacent
//     lines (the stored order is by column length); the basis is also written in natural order for the callers;
//   * alpha, beta and the bookkeeping of cubic.py:98-109 are replicated in the registers of every CTA: after a
//     grid barrier each CTA adds the 148 per-CTA pieces in the same fixed order, so all take the same branch.
//
// Grid barrier: arrivals are counted in global memory (release fence, one atomic per CTA, acquire poll), all CTAs
// are co-resident by construction (grid = number of SMs, one CTA per SM by shared-memory size).
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "kcrn_device.cuh"
#include "kcrn_internal.cuh"

namespace kcrn {

// ----------------------------------------------------------------------------- tile planning (host)
// Balanced tiles: tile k takes the next (remaining entries) / (remaining tiles) entries.  A short segment that
// crosses the boundary goes to the side that holds most of it; a segment longer than kPSplit is cut exactly at the
// boundary into pieces (first piece = tail of a tile, middle pieces = whole tiles, last piece = head of a tile).
// Empty segments belong to the tile they are met in.  Returns false when the problem does not fit `grid` tiles.
bool plan_persist_tiles(const std::vector<int32_t> &ptr, int64_t nseg, int64_t nnz, int grid, std::vector<PTile> &tiles,
                        std::vector<int4> &longtab, int &npartial) {
    tiles.clear();
    longtab.clear();
    npartial = 0;
    if (nnz <= 0 || nseg <= 0 || grid < 1) return false;
    int64_t s = 0, off = 0;          // next segment; entries of it already given to earlier tiles
    int cur_long = -1, cur_piece = 0;
    while (s < nseg) {
        if ((int)tiles.size() >= grid) return false;
        const int64_t e0 = ptr[s] + off;
        const int64_t left = grid - (int64_t)tiles.size();
        const int64_t share = (nnz - e0 + left - 1) / left;
        if (share > kPCap - kPSplit) return false;
        const int64_t target = e0 + share;
        PTile t{};
        t.e0 = (int)e0;
        t.seg0 = (int)s;
        t.head_long = t.tail_long = -1;
        if (off > 0) { t.head_long = cur_long; t.head_piece = cur_piece; t.pad[0] = longtab[cur_long].x + cur_piece; }
        int64_t cur = e0;
        int ns = 0;
        while (s < nseg && ns < kPSegs) {
            const int64_t send = ptr[s + 1];
            if (send <= target) {                      // the rest of the segment fits (empty segments included)
                if (off > 0) {                         // ... and closes a cut segment: this is its last piece
                    longtab[cur_long].y = cur_piece + 1;
                    npartial += 1;
                    off = 0;
                    cur_long = -1;
                }
                cur = send;
                ++ns;
                ++s;
                continue;
            }
            const int64_t total_len = (int64_t)ptr[s + 1] - ptr[s];
            if (total_len > kPSplit) {
                if (target > cur) {                    // cut: this tile takes [cur, target)
                    if (off == 0) {
                        longtab.push_back(make_int4(npartial, 0, (int)s, 0));
                        cur_long = (int)longtab.size() - 1;
                        cur_piece = 0;
                    }
                    t.tail_long = cur_long;
                    t.tail_piece = cur_piece;
                    t.pad[1] = longtab[cur_long].x + cur_piece;
                    npartial += 1;
                    cur_piece += 1;
                    off = target - ptr[s];
                    cur = target;
                    ++ns;
                }
                break;
            }
            // short segment across the boundary: whole, on the side that holds most of it
            if (ns == 0 || (send - target) <= (target - std::max<int64_t>(cur, ptr[s]))) {
                cur = send;
                ++ns;
                ++s;
            }
            break;
        }
        t.nk = (int)(cur - e0);
        t.nseg = ns;
        if (t.nk > kPCap || ns == 0) return false;
        // threads per short segment: cheapest of 1, 2, .. 32 under the model of the per-pass kernels
        int maxlen = 0, nshort = 0;
        for (int i = 0; i < ns; ++i) {
            const int64_t a = std::max<int64_t>(ptr[t.seg0 + i], e0), b = std::min<int64_t>(ptr[t.seg0 + i + 1], cur);
            const int len = (int)std::max<int64_t>(0, b - a);
            if (len <= kPBig) { maxlen = std::max(maxlen, len); ++nshort; }
        }
        int best = 0;
        long best_cost = -1;
        for (int lg = 0; lg <= 5; ++lg) {
            const int G = 1 << lg, groups = kPThreads >> lg;
            const long rounds = (ns + groups - 1) / groups;
            const long cost = rounds * ((maxlen + G - 1) / G + 2 * lg + 3);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = lg; }
        }
        t.lg = best;
        tiles.push_back(t);
    }
    return (int)tiles.size() <= grid;
}

namespace {

// ----------------------------------------------------------------------------- PTX helpers
// (mbarrier / bulk-copy helpers: kcrn_device.cuh)
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// vectors other CTAs wrote before the last exchange: plain (coherent) loads, L1 allocation allowed -- the acquire
// fence after the exchange invalidated this SM's L1, and nobody writes them again before the next exchange
__device__ __forceinline__ double ld_vec(const double *p) {
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// Exchange slots: one 16-byte word {value.lo, epoch, value.hi, epoch} per published double, the layout of NCCL's
// low-latency protocol -- the value carries its own flag, so "has everybody arrived" and "what did they leave" are ONE
// round trip to L2.  Two halves by the parity of the epoch: a CTA can run at most one exchange ahead of the slowest.
__device__ __forceinline__ void ll_store(uint4 *slot, double v, unsigned epoch) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(slot), "r"((unsigned)__double2loint(v)), "r"(epoch),
                 "r"((unsigned)__double2hiint(v)), "r"(epoch)
                 : "memory");
}
__device__ __forceinline__ double ll_wait(const uint4 *slot, unsigned epoch) {
    unsigned lo, f0, hi, f1;
    unsigned long long t0 = 0;
    for (unsigned spin = 0;; ++spin) {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(slot) : "memory");
        if (f0 == epoch && f1 == epoch) break;
        if ((spin & 0xfffffu) == 0xfffffu) {           // a CTA that never arrives must not hang the box
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 10000000000ull) __trap();
        }
    }
    return __hiloint2double((int)hi, (int)lo);
}

// two slots at once (both loads in flight together)
__device__ __forceinline__ void ll_wait2(const uint4 *s0, const uint4 *s1, unsigned epoch, double &v0, double &v1) {
    unsigned a[4], b[4];
    unsigned long long t0 = 0;
    for (unsigned spin = 0;; ++spin) {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "l"(s0) : "memory");
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]) : "l"(s1) : "memory");
        if (a[1] == epoch && a[3] == epoch && b[1] == epoch && b[3] == epoch) break;
        if ((spin & 0xfffffu) == 0xfffffu) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 10000000000ull) __trap();
        }
    }
    v0 = __hiloint2double((int)a[2], (int)a[0]);
    v1 = __hiloint2double((int)b[2], (int)b[0]);
}

struct PMatView {
    const PTile *tiles;
    int ntiles;
    const uint32_t *word;      // [ntiles][kPCap] gather-ordered words
    const double *tval;        // [ntiles][kPCap] values in the same order
    const int32_t *ptr;
    const int4 *longtab;       // cut segments {first piece slot, pieces, stored segment, 0}
    int nlong, npieces;
};

struct PersistParams {
    PMatView R, C;
    const int32_t *segid;      // stored position of the CSC copy -> original column (null: identity)
    const double *D;
    double *t;
    const double *g;           // gradient, natural order
    double *gp, *wp, *Vp;      // stored column order: gradient, w', basis [m][d]
    double *V;                 // basis in natural order [m][d]
    double *alphas, *bet;
    LanczosCtl *ctl;
    int32_t *info;
    int64_t n, d;
    int m;
    double n_samples, inv_n, l2;
    uint4 *cta_slots;          // [2 parities][2 values][grid]
    uint4 *piece_slots;        // [2 parities][piece_cap]
    int piece_cap;
    unsigned *epoch_base;      // exchanges done by earlier launches
    int flags;                 // experiments (KCRN_PERSIST_FLAGS): 1 = gathers bypass L1, no acquire fence; 2 = no natural-order basis store
    long long *trace;          // KCRN_PERSIST_TRACE=1: [grid][kPTraceCap] SM clock at the marks of thread 0 (null: off)
};

constexpr int kPTraceCap = 512;
#define PTRACE()                                                              \
    do {                                                                      \
        if (P.trace && threadIdx.x == 0 && ntrace < kPTraceCap)               \
            P.trace[(size_t)blockIdx.x * kPTraceCap + ntrace++] = clock64();  \
    } while (0)

constexpr int kPWarps = kPThreads / 32;
constexpr int kPItems = kPCap / kPThreads;     // 12 entries per thread
constexpr int kPCutCap = 1024;                 // cut segments per matrix (one thread each in the fix-up)
constexpr size_t kPSmemWord = sizeof(uint32_t) * kPCap;
constexpr size_t kPSmemVal = sizeof(double) * kPCap;
constexpr size_t kPSmemSum = sizeof(double) * kPSegs;
constexpr size_t kPSmemPtr = (sizeof(uint16_t) * (kPSegs + 1) + 15) & ~(size_t)15;
constexpr size_t kPSmemOps = sizeof(double) * (2048 + 2 * kPCutCap);      // epilogue operands staged by cp.async
constexpr size_t kPSmemBytes = kPSmemWord + kPSmemVal + kPSmemSum + 2 * kPSmemPtr + kPSmemOps;

struct PShared {
    uint32_t *word;
    double *val;       // values of the tile in gather order, then the products in natural order
    double *sum;       // per-segment sums; between passes: piece sums [0, 2048) and cut-segment shares [2048, 3072)
    uint16_t *ptrR, *ptrC;
};

// CTA sum in fixed order; thread 0 gets the total
__device__ __forceinline__ double cta_total(double v, double *sred) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
    __syncthreads();
    double tot = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kPWarps; ++w) tot += sred[w];
    __syncthreads();
    return tot;
}

// bulk copies of one tile: words, then values (each on its own mbarrier); thread 0 only
__device__ __forceinline__ void request_words(const PMatView &M, const PTile &t, int tile, const PShared &S, uint64_t *bar) {
    const unsigned bytes = ((unsigned)t.nk * 4u + 15u) & ~15u;
    fence_proxy_async();
    mbar_expect_tx(bar, bytes);
    bulk_g2s(S.word, M.word + (size_t)tile * kPCap, bytes, bar);
}
__device__ __forceinline__ void request_values(const PMatView &M, const PTile &t, int tile, const PShared &S, uint64_t *bar) {
    const unsigned bytes = ((unsigned)t.nk * 8u + 15u) & ~15u;
    fence_proxy_async();
    mbar_expect_tx(bar, bytes);
    bulk_g2s(S.val, M.tval + (size_t)tile * kPCap, bytes, bar);
}

// tile-local segment pointers (clipped to the tile: pieces of cut segments start at 0 / end at nk)
__device__ __forceinline__ void fill_local_ptr(const PMatView &M, const PTile &t, uint16_t *sp) {
    for (int i = threadIdx.x; i <= t.nseg; i += kPThreads) {
        int p = __ldg(M.ptr + t.seg0 + i) - t.e0;
        p = max(0, min(t.nk, p));
        sp[i] = (uint16_t)p;
    }
}

// All gathers of the tile are issued before anything is consumed (one L2 round trip per pass, not one per batch)
__device__ __forceinline__ void issue_gathers(const uint32_t *sw, int nk, const double *src, double (&x)[kPItems], bool cg = false) {
    if (cg) {
#pragma unroll
        for (int u = 0; u < kPItems; ++u) {
            const int i = u * kPThreads + (int)threadIdx.x;
            x[u] = i < nk ? __ldcg(src + (sw[i] >> kPSlotBits)) : 0.0;
        }
        return;
    }
#pragma unroll
    for (int u = 0; u < kPItems; ++u) {
        const int i = u * kPThreads + (int)threadIdx.x;
        x[u] = i < nk ? ld_vec(src + (sw[i] >> kPSlotBits)) : 0.0;
    }
}

// products: values are read in gather order (conflict-free), the products return to their natural positions after
// every thread has consumed its values (the product buffer IS the value buffer)
template <bool kDivide>
__device__ __forceinline__ void multiply_scatter(const uint32_t *sw, double *sv, int nk, double beta, double rinv,
                                                 double (&x)[kPItems]) {
#pragma unroll
    for (int u = 0; u < kPItems; ++u) {
        const int i = u * kPThreads + (int)threadIdx.x;
        if (i < nk) {
            double xv = x[u];
            if constexpr (kDivide) xv = quotient(xv, beta, rinv);
            x[u] = sv[i] * xv;
        }
        if ((u & 3) == 3) asm volatile("" ::: "memory");      // keep the value loads from piling up in registers
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kPItems; ++u) {
        const int i = u * kPThreads + (int)threadIdx.x;
        if (i < nk) sv[sw[i] & ((1u << kPSlotBits) - 1u)] = x[u];
        if ((u & 3) == 3) asm volatile("" ::: "memory");
    }
    __syncthreads();
}

// per-segment sums of the products: groups of 2^lg threads walk short segments left to right; segments longer than
// kPBig are summed by the whole CTA (fixed order: thread-strided, warp tree, warps in order)
__device__ __forceinline__ void segment_sums(const PTile &t, const uint16_t *sp, const double *sv, double *ssum,
                                             double *sred, int *s_big, int *s_nbig) {
    const int tid = threadIdx.x, lg = t.lg, G = 1 << lg, groups = kPThreads >> lg, gid = tid >> lg, gl = tid & (G - 1);
    for (int sb = 0; sb < t.nseg; sb += groups) {
        const int sg = sb + gid;
        int a = 0, b = 0;
        if (sg < t.nseg) { a = sp[sg]; b = sp[sg + 1]; }
        const bool big = b - a > kPBig;
        double sum = 0.0;
        if (!big)
            for (int i = a + gl; i < b; i += G) sum += sv[i];
        for (int o = G >> 1; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (sg < t.nseg && gl == 0) {
            if (big) s_big[atomicAdd(s_nbig, 1)] = sg;
            else ssum[sg] = sum;
        }
    }
    __syncthreads();
    const int nbig = *s_nbig;
    for (int k = 0; k < nbig; ++k) {
        const int sg = s_big[k], a = sp[sg], b = sp[sg + 1];
        double s = 0.0;
        for (int i = a + tid; i < b; i += kPThreads) s += sv[i];
        const double tot = cta_total(s, sred);
        if (tid == 0) ssum[sg] = tot;
    }
    __syncthreads();
}

// local segment sg of tile t as a piece of a cut segment: its slot among the matrix's piece sums, or -1
__device__ __forceinline__ int piece_slot(const PTile &t, int sg) {
    if (sg == 0 && t.head_long >= 0) return t.pad[0];
    if (sg == t.nseg - 1 && t.tail_long >= 0) return t.pad[1];
    return -1;
}

// The exchange that ends a pass.  Before: every CTA has stored its part of the pass's output vector and the sums of
// its pieces of cut segments (ll_store into piece_slots).  Here: publish up to two per-CTA numbers, wait for
// everybody's, and return their totals -- added in CTA order by every CTA, so all get the same bits.  The piece sums
// of the cut segments land in S.sum[0 .. npieces).  On return the output vector of the pass is complete and visible
// (except the cut segments, which the caller finishes from the piece sums, identically in every CTA).
template <int NV>
__device__ __forceinline__ void exchange(const PersistParams &P, unsigned &epoch, double mine0, double mine1, int npieces,
                                         double *spiece, double *sred, double &tot0, double &tot1, int &ntrace,
                                         const PMatView *reqM = nullptr, const PTile *reqT = nullptr, const PShared *reqS = nullptr,
                                         uint64_t *reqBar = nullptr) {
    const int tid = threadIdx.x, G = gridDim.x;
    epoch += 1;
    uint4 *cs = P.cta_slots + (size_t)(epoch & 1u) * 2 * G;
    const uint4 *ps = P.piece_slots + (size_t)(epoch & 1u) * P.piece_cap;
    __syncthreads();                                   // all stores of the pass have been issued by this CTA
    PTRACE();                                          // x+0: CTA sums
    if (tid == 0) {
        fence_gpu();                                   // release
        ll_store(cs + blockIdx.x, mine0, epoch);
        if (NV > 1) ll_store(cs + G + blockIdx.x, mine1, epoch);
        // the next tile's values are requested only now: a bulk copy in flight would hold up the fence above
        if (reqM) request_values(*reqM, *reqT, blockIdx.x, *reqS, reqBar);
    }
    PTRACE();                                          // x+1: release fence, publish
    double a0 = 0.0, a1 = 0.0;
    if (tid < G) {
        if (NV > 1) ll_wait2(cs + tid, cs + G + tid, epoch, a0, a1);
        else a0 = ll_wait(cs + tid, epoch);
    }
    for (int q = tid - G; q < npieces; q += kPThreads - G)
        if (q >= 0) spiece[q] = ll_wait(ps + q, epoch);
    // acquire: ONE fence after every flag has been seen (it invalidates this SM's L1; a fence per polling thread
    // cost 7 us per exchange)
    __syncthreads();
    PTRACE();                                          // x+2: everybody's numbers have arrived
    if (tid == 0 && !(P.flags & 1)) fence_gpu();
    PTRACE();                                          // x+3: acquire fence
    a0 = warp_sum(a0);
    if (NV > 1) a1 = warp_sum(a1);
    if ((tid & 31) == 0) { sred[tid >> 5] = a0; if (NV > 1) sred[kPWarps + (tid >> 5)] = a1; }
    __syncthreads();
    tot0 = tot1 = 0.0;
    const int nw = (G + 31) >> 5;
    for (int w = 0; w < nw; ++w) { tot0 += sred[w]; if (NV > 1) tot1 += sred[kPWarps + w]; }
    __syncthreads();
}

// sum of the cut segments' shares S[0 .. n) in a fixed order, the same in every thread
__device__ __forceinline__ double cut_total(const double *s, int n, double *sred) {
    if (n == 0) return 0.0;
    double x = 0.0;
    for (int i = threadIdx.x; i < n; i += kPThreads) x += s[i];
    x = warp_sum(x);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = x;
    __syncthreads();
    double tot = 0.0;
    const int nw = min(kPWarps, (n + 31) >> 5);
    for (int w = 0; w < nw; ++w) tot += sred[w];
    __syncthreads();
    return tot;
}

__global__ void __launch_bounds__(kPThreads, 1) persist_lanczos_kernel(const PersistParams P) {
    extern __shared__ __align__(128) unsigned char psm[];
    __shared__ uint64_t barW, barV;
    __shared__ double sred[2 * kPWarps];
    __shared__ int s_big[16], s_nbig;
    __shared__ PTile s_tiles[2];               // descriptors and recurrence state in shared memory: registers are for
    __shared__ LanczosCtl s_st;                // the gathers in flight
    PShared S;
    S.word = reinterpret_cast<uint32_t *>(psm);
    S.val = reinterpret_cast<double *>(psm + kPSmemWord);
    S.sum = reinterpret_cast<double *>(psm + kPSmemWord + kPSmemVal);
    S.ptrR = reinterpret_cast<uint16_t *>(psm + kPSmemWord + kPSmemVal + kPSmemSum);
    S.ptrC = reinterpret_cast<uint16_t *>(psm + kPSmemWord + kPSmemVal + kPSmemSum + kPSmemPtr);
    // epilogue operands, fetched with cp.async at the start of a pass (no registers while the gathers fly):
    // [0] first segment round, operand 0; [1] the same, operand 1; [2] cut segment of this thread: operands 0, 1 interleaved
    double *const sop = reinterpret_cast<double *>(psm + kPSmemWord + kPSmemVal + kPSmemSum + 2 * kPSmemPtr);
    double *const spiece = S.sum, *const sshare = S.sum + 2048;
    const int tid = threadIdx.x, cta = blockIdx.x, G = gridDim.x;
    const bool scribe = cta == 0 && tid == 0;
    unsigned epoch = *P.epoch_base;            // read before this CTA publishes anything (see the end of the kernel)
    int ntrace = 0;
    PTRACE();                                  // mark 0: kernel start

    // my two tiles (constant for the whole kernel), their local pointers, the first stream request
    const bool haveR = cta < P.R.ntiles, haveC = cta < P.C.ntiles;
    if (tid < 24) {
        const int which = tid / 12, f = tid % 12;
        int v = 0;
        if (which == 0 ? haveR : haveC) v = reinterpret_cast<const int *>((which == 0 ? P.R.tiles : P.C.tiles) + cta)[f];
        reinterpret_cast<int *>(s_tiles + which)[f] = v;
    }
    if (tid == 0) {
        mbar_init(&barW, 1);
        mbar_init(&barV, 1);
        fence_mbar_init();
        s_nbig = 0;
    }
    __syncthreads();
    const PTile &tr = s_tiles[0], &tc = s_tiles[1];
    const bool streamR = haveR && tr.nk > 0, streamC = haveC && tc.nk > 0;
    if (haveR) fill_local_ptr(P.R, tr, S.ptrR);
    if (haveC) fill_local_ptr(P.C, tc, S.ptrC);
    if (tid == 0 && streamR) {
        request_words(P.R, tr, cta, S, &barW);
        request_values(P.R, tr, cta, S, &barV);
    }
    unsigned phase = 0;
    // the cut segment this thread finishes after an exchange: {first piece slot | pieces << 16, stored segment}
    int cutR_sp = 0, cutR_seg = 0, cutC_sp = 0, cutC_seg = 0;
    if (tid < P.R.nlong) { const int4 lt = __ldg(P.R.longtab + tid); cutR_sp = lt.x | (lt.y << 16); cutR_seg = lt.z; }
    if (tid < P.C.nlong) { const int4 lt = __ldg(P.C.longtab + tid); cutC_sp = lt.x | (lt.y << 16); cutC_seg = lt.z; }

    // ---- start: the gradient in stored column order, |g|
    {
        double loc = 0.0;
        for (int64_t p = (int64_t)cta * kPThreads + tid; p < P.d; p += (int64_t)G * kPThreads) {
            const double x = __ldg(P.g + (P.segid ? __ldg(P.segid + p) : (int)p));
            P.gp[p] = x;
            loc = fma(x, x, loc);
        }
        const double mine = cta_total(loc, sred);
        if (cta == 0)
            for (int i = tid; i < P.m; i += kPThreads) { P.alphas[i] = 0.0; P.bet[i + 1] = 0.0; }
        double tot, unused;
        exchange<1>(P, epoch, mine, 0.0, 0, spiece, sred, tot, unused, ntrace);
        const double gnorm = sqrt(tot);
        if (tid == 0) {
            LanczosCtl st;
            st.col = 0; st.redo = 0; st.final_ = (P.m == 1); st.done = 0; st.m_eff = P.m; st.nvalid = 1;
            st.store_idx = P.m - 1; st.m = P.m; st.broke = 0; st.hvps = 0; st.last_beta = 0.0;
            st.cur_beta = gnorm; st.cur_alpha = 0.0;
            s_st = st;
            if (cta == 0) P.bet[0] = gnorm;
        }
        __syncthreads();
    }
    PTRACE();                                  // mark 1: |g| known

    for (;;) {
        // ================================================== row pass: t = D (A v), alpha = v'Hv
        const int col = s_st.col;
        const bool redo = s_st.redo != 0, final_ = s_st.final_ != 0;
        const double *src = redo ? P.Vp + (size_t)col * P.d : (col == 0 ? P.gp : P.wp);
        const double beta = redo ? 1.0 : s_st.cur_beta, rinv = 1.0 / beta;
        const bool more = !final_;             // a column pass follows
        double acc = 0.0, vsq = 0.0;
        {
            double x[kPItems];
            // loads that do not depend on the tile: the element of the basis column this thread normalises, the
            // epilogue operands (straight to shared memory)
            const int64_t pmine = (int64_t)cta * kPThreads + tid;
            const bool side = (!redo || P.l2 != 0.0) && pmine < P.d;
            const double vsrc = side ? ld_vec(src + pmine) : 0.0;
            if (haveR && tid < tr.nseg) cp_async_8(sop + tid, P.D + tr.seg0 + tid);
            if (tid < P.R.nlong) cp_async_8(sop + 2048 + 2 * tid, P.D + cutR_seg);
            if (streamR) {
                mbar_wait(&barW, phase);
                PTRACE();                      // row +0: words have landed
                issue_gathers(S.word, tr.nk, src, x, P.flags & 1);
                mbar_wait(&barV, phase);
                multiply_scatter<true>(S.word, S.val, tr.nk, beta, rinv, x);
                phase ^= 1;
            } else PTRACE();
            PTRACE();                          // row +1: products in place
            // basis column `col` (both orders) and |v|^2 for the l2 part of alpha
            if (side) {
                double *vp = P.Vp + (size_t)col * P.d, *vn = P.V + (size_t)col * P.d;
                for (int64_t p = pmine; p < P.d; p += (int64_t)G * kPThreads) {
                    const double vc = quotient(p == pmine ? vsrc : ld_vec(src + p), beta, rinv);
                    if (!redo) {
                        vp[p] = vc;
                        if (!(P.flags & 2)) vn[P.segid ? __ldg(P.segid + p) : (int)p] = vc;
                    }
                    vsq = fma(vc, vc, vsq);
                }
            }
        }
        if (tid == 0 && more && streamC) request_words(P.C, tc, cta, S, &barW);       // the word buffer is free
        cp_async_wait_all();
        PTRACE();                              // row +2: basis column stored
        if (haveR) {
            segment_sums(tr, S.ptrR, S.val, S.sum, sred, s_big, &s_nbig);
            PTRACE();                          // row +3: segment sums
            if (tid == 0) {
                s_nbig = 0;
                if (more && streamC && !(P.flags & 4)) request_values(P.C, tc, cta, S, &barV);        // the value buffer is free
            }
            const unsigned next = epoch + 1;
            uint4 *ps = P.piece_slots + (size_t)(next & 1u) * P.piece_cap;
            for (int sg = tid; sg < tr.nseg; sg += kPThreads) {
                const double sum = S.sum[sg];
                const int slot = piece_slot(tr, sg);
                if (slot >= 0) {
                    ll_store(ps + slot, sum, next);
                } else {
                    const double dv = sg == tid ? sop[tid] : __ldg(P.D + tr.seg0 + sg);
                    const double tv = __dmul_rn(dv, sum);
                    P.t[tr.seg0 + sg] = tv;
                    acc += __dmul_rn(tv, sum);
                }
            }
        } else if (tid == 0 && more && streamC && !(P.flags & 4)) {
            request_values(P.C, tc, cta, S, &barV);
        }
        if (!haveR) PTRACE();
        __syncthreads();                       // S.sum is free for the exchange
        PTRACE();                              // row +4: epilogue
        {
            const double m0 = cta_total(acc, sred), m1 = cta_total(vsq, sred);
            double s1, s2;
            {
                const bool late = (P.flags & 4) && more && streamC;
                exchange<2>(P, epoch, m0, m1, P.R.npieces, spiece, sred, s1, s2, ntrace, late ? &P.C : nullptr, &tc, &S, &barV);
            }
            PTRACE();                          // row +5: exchange
            // cut rows: every CTA adds the pieces in piece order and applies the epilogue (same bits everywhere)
            if (P.R.nlong > 0) {
                if (tid < P.R.nlong) {
                    double sum = 0.0;
                    const int s0 = cutR_sp & 0xffff, np = cutR_sp >> 16;
                    for (int i = 0; i < np; ++i) sum += spiece[s0 + i];
                    const double tv = __dmul_rn(sop[2048 + 2 * tid], sum);
                    P.t[cutR_seg] = tv;
                    sshare[tid] = __dmul_rn(tv, sum);
                }
                __syncthreads();
                s1 += cut_total(sshare, P.R.nlong, sred);
            }
            const double alpha = __dadd_rn(quotient(s1, P.n_samples, P.inv_n), __dmul_rn(P.l2, s2));
            if (tid == 0) {
                s_st.cur_alpha = alpha;
                s_st.hvps += 1;
                if (cta == 0) P.alphas[final_ ? s_st.store_idx : col] = alpha;
                if (final_) s_st.done = 1;     // cubic.py:109: the last alpha needs no further vector
            }
            __syncthreads();
            PTRACE();                          // row +6: alpha
            if (final_) break;
        }

        // ================================================== column pass: w' = A't/n + l2 v - beta v_prev - alpha v
        acc = 0.0;
        const double *vcol = P.Vp + (size_t)col * P.d, *vpre = P.Vp + (size_t)(col > 0 ? col - 1 : 0) * P.d;
        const bool subtract = col > 0;
        const double beta_prev = s_st.cur_beta, alpha = s_st.cur_alpha;
        {
            double x[kPItems];
            // operands: first segment round of the tile, and the cut column this thread finishes after the exchange
            if (haveC && tid < tc.nseg) {
                cp_async_8(sop + tid, vcol + tc.seg0 + tid);
                if (subtract) cp_async_8(sop + 1024 + tid, vpre + tc.seg0 + tid);
            }
            if (tid < P.C.nlong) {
                cp_async_8(sop + 2048 + 2 * tid, vcol + cutC_seg);
                if (subtract) cp_async_8(sop + 2048 + 2 * tid + 1, vpre + cutC_seg);
            }
            if (streamC) {
                mbar_wait(&barW, phase);
                PTRACE();                      // col +0: words have landed
                issue_gathers(S.word, tc.nk, P.t, x, P.flags & 1);
                mbar_wait(&barV, phase);
                multiply_scatter<false>(S.word, S.val, tc.nk, 1.0, 1.0, x);
                phase ^= 1;
            } else PTRACE();
            PTRACE();                          // col +1: products in place
        }
        if (tid == 0 && streamR) request_words(P.R, tr, cta, S, &barW);      // a row pass always follows
        cp_async_wait_all();
        if (haveC) {
            segment_sums(tc, S.ptrC, S.val, S.sum, sred, s_big, &s_nbig);
            PTRACE();                          // col +2: segment sums
            if (tid == 0) {
                s_nbig = 0;
                if (streamR && !(P.flags & 4)) request_values(P.R, tr, cta, S, &barV);
            }
            const unsigned next = epoch + 1;
            uint4 *ps = P.piece_slots + (size_t)(next & 1u) * P.piece_cap;
            for (int sg = tid; sg < tc.nseg; sg += kPThreads) {
                const double sum = S.sum[sg];
                const int slot = piece_slot(tc, sg);
                if (slot >= 0) {
                    ll_store(ps + slot, sum, next);
                } else {
                    const double v0 = sg == tid ? sop[tid] : ld_vec(vcol + tc.seg0 + sg);
                    const double v1 = !subtract ? 0.0 : (sg == tid ? sop[1024 + tid] : ld_vec(vpre + tc.seg0 + sg));
                    double hv = __dadd_rn(quotient(sum, P.n_samples, P.inv_n), __dmul_rn(P.l2, v0));
                    if (subtract) hv = __dsub_rn(hv, __dmul_rn(beta_prev, v1));
                    hv = __dsub_rn(hv, __dmul_rn(alpha, v0));
                    P.wp[tc.seg0 + sg] = hv;
                    acc += __dmul_rn(hv, hv);
                }
            }
        } else if (tid == 0 && streamR && !(P.flags & 4)) {
            request_values(P.R, tr, cta, S, &barV);
        }
        if (!haveC) PTRACE();
        __syncthreads();
        PTRACE();                              // col +3: epilogue
        {
            const double m2 = cta_total(acc, sred);
            double bsq, unused;
            {
                const bool late = (P.flags & 4) && streamR;
                exchange<1>(P, epoch, m2, 0.0, P.C.npieces, spiece, sred, bsq, unused, ntrace, late ? &P.R : nullptr, &tr, &S, &barV);
            }
            PTRACE();                          // col +4: exchange
            if (P.C.nlong > 0) {
                if (tid < P.C.nlong) {
                    double sum = 0.0;
                    const int s0 = cutC_sp & 0xffff, np = cutC_sp >> 16;
                    for (int i = 0; i < np; ++i) sum += spiece[s0 + i];
                    const double c0 = sop[2048 + 2 * tid], c1 = subtract ? sop[2048 + 2 * tid + 1] : 0.0;
                    double hv = __dadd_rn(quotient(sum, P.n_samples, P.inv_n), __dmul_rn(P.l2, c0));
                    if (subtract) hv = __dsub_rn(hv, __dmul_rn(beta_prev, c1));
                    hv = __dsub_rn(hv, __dmul_rn(alpha, c0));
                    P.wp[cutC_seg] = hv;               // every CTA writes the same bits; each reads back its own
                    sshare[tid] = __dmul_rn(hv, hv);
                }
                __syncthreads();
                bsq += cut_total(sshare, P.C.nlong, sred);
            }
            if (tid == 0) lanczos_advance(&s_st, cta == 0 ? P.bet : nullptr, sqrt(bsq));       // cubic.py:98-109
            __syncthreads();
            PTRACE();                          // col +5: beta
        }
    }

    if (scribe) {
        const LanczosCtl st = s_st;
        P.bet[st.m] = st.last_beta;
        P.info[0] = st.m_eff; P.info[1] = st.nvalid; P.info[2] = st.broke; P.info[3] = st.hvps;
        P.info[4] = st.col; P.info[5] = st.done; P.info[6] = st.redo; P.info[7] = st.final_;
        *P.ctl = st;
        *P.epoch_base = epoch;                 // every CTA read it before the first exchange of this launch completed
    }
}

// ----------------------------------------------------------------------------- set-up kernels
// One CTA per tile: words `gather index << 14 | position` sorted by gather index (bitonic network in shared memory),
// and the values in the same order
__global__ void __launch_bounds__(kPThreads) persist_words_kernel(const PTile *__restrict__ tiles,
                                                                 const int32_t *__restrict__ idx,
                                                                 const double *__restrict__ val,
                                                                 const int32_t *__restrict__ renumber,
                                                                 uint32_t *__restrict__ word, double *__restrict__ tval) {
    extern __shared__ uint32_t key[];          // 1 << kPSlotBits
    const PTile t = tiles[blockIdx.x];
    const int tid = threadIdx.x, nk = t.nk;
    int np = 2;
    while (np < nk) np <<= 1;
    for (int i = tid; i < np; i += kPThreads) {
        uint32_t k = 0xffffffffu;
        if (i < nk) {
            int c = idx[t.e0 + i];
            if (renumber) c = renumber[c];
            k = ((uint32_t)c << kPSlotBits) | (uint32_t)i;
        }
        key[i] = k;
    }
    __syncthreads();
    for (int k = 2; k <= np; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int q = tid; q < (np >> 1); q += kPThreads) {
                const int i = ((q & ~(j - 1)) << 1) | (q & (j - 1)), l = i | j;
                const uint32_t a = key[i], b = key[l];
                if ((a > b) == ((i & k) == 0)) { key[i] = b; key[l] = a; }
            }
            __syncthreads();
        }
    uint32_t *out = word + (size_t)blockIdx.x * kPCap;
    double *vout = tval + (size_t)blockIdx.x * kPCap;
    for (int i = tid; i < kPCap; i += kPThreads) {
        const uint32_t w = i < nk ? key[i] : 0xffffffffu;
        out[i] = w;
        vout[i] = i < nk ? val[t.e0 + (int)(w & ((1u << kPSlotBits) - 1u))] : 0.0;
    }
}

int build_persist_mat(const SegMat &A, const int32_t *renumber, int grid, PersistMat *M, cudaStream_t s, bool *ok, bool verbose) {
    *ok = false;
    std::vector<int32_t> ptr((size_t)A.nseg + 1);
    KCRN_CUDA(cudaMemcpyAsync(ptr.data(), A.ptr, sizeof(int32_t) * (A.nseg + 1), cudaMemcpyDeviceToHost, s));
    KCRN_CUDA(cudaStreamSynchronize(s));
    std::vector<PTile> tiles;
    std::vector<int4> longtab;
    int npartial = 0;
    if (!plan_persist_tiles(ptr, A.nseg, A.nnz, grid, tiles, longtab, npartial)) return KCRN_OK;
    if ((int)longtab.size() > kPCutCap || npartial > 2048) return KCRN_OK;
    M->ntiles = (int)tiles.size();
    M->nlong = (int)longtab.size();
    M->npartial = npartial;
    KCRN_CUDA(cudaMalloc(&M->tiles, sizeof(PTile) * tiles.size()));
    KCRN_CUDA(cudaMalloc(&M->word, sizeof(uint32_t) * (size_t)kPCap * tiles.size()));
    KCRN_CUDA(cudaMalloc(&M->tval, sizeof(double) * (size_t)kPCap * tiles.size()));
    KCRN_CUDA(cudaMalloc(&M->longtab, sizeof(int4) * std::max<size_t>(longtab.size(), 1)));
    KCRN_CUDA(cudaMemcpyAsync(M->tiles, tiles.data(), sizeof(PTile) * tiles.size(), cudaMemcpyHostToDevice, s));
    if (!longtab.empty())
        KCRN_CUDA(cudaMemcpyAsync(M->longtab, longtab.data(), sizeof(int4) * longtab.size(), cudaMemcpyHostToDevice, s));
    const size_t smem = sizeof(uint32_t) << kPSlotBits;
    KCRN_CUDA(cudaFuncSetAttribute(persist_words_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    persist_words_kernel<<<(unsigned)tiles.size(), kPThreads, smem, s>>>(M->tiles, A.idx, A.val, renumber, M->word, M->tval);
    KCRN_LAUNCH_CHECK();
    KCRN_CUDA(cudaStreamSynchronize(s));
    if (verbose)
        fprintf(stderr, "kcrn: persistent tiles: %lld segments, %lld non-zeros -> %d tiles (%d cut segments, %d pieces)\n",
                (long long)A.nseg, (long long)A.nnz, M->ntiles, M->nlong, npartial);
    *ok = true;
    return KCRN_OK;
}

void free_persist_mat(PersistMat *M) {
    cudaFree(M->tiles); cudaFree(M->word); cudaFree(M->tval); cudaFree(M->longtab);
    *M = PersistMat();
}

PMatView view_of(const PersistMat &M, const SegMat &A) {
    PMatView v{};
    v.tiles = M.tiles; v.ntiles = M.ntiles; v.word = M.word; v.tval = M.tval; v.ptr = A.ptr; v.longtab = M.longtab;
    v.nlong = M.nlong; v.npieces = M.npartial;
    return v;
}

}  // namespace

// Called at the end of kcrn_create.  The persistent path serves single-GPU handles whose two copies of A fit one
// tile per SM and whose gather lengths fit the word; everything else keeps the graph of per-pass kernels.
int persist_build(kcrn_problem *h) {
    h->persist_ok = false;
    // Opt-in (KCRN_PERSIST=1): measured on a B200 it is correct but slower than the graph of per-pass kernels at the
    // rcv1 shape (27 vs 20.7 us per HVP; DESIGN.md section 3.4 has the per-phase trace and why)
    if (!h->sw.persist || h->sharded() || h->nnz <= 0) return KCRN_OK;
    if (h->d >= kPMaxIndex - 1 || h->n >= kPMaxIndex - 1) return KCRN_OK;
    int dev = 0, sms = 0, per_sm = 0;
    KCRN_CUDA(cudaGetDevice(&dev));
    KCRN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    KCRN_CUDA(cudaFuncSetAttribute(persist_lanczos_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPSmemBytes));
    KCRN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, persist_lanczos_kernel, kPThreads, kPSmemBytes));
    if (per_sm < 1 || sms < 1) return KCRN_OK;
    const int grid = sms;                      // one CTA per SM: the barrier needs all of them resident
    bool okr = false, okc = false;
    KCRN_TRY(build_persist_mat(h->csr, h->csc.pos, grid, &h->prow, h->stream, &okr, h->sw.verbose));     // rows gather stored positions
    if (okr) KCRN_TRY(build_persist_mat(h->csc, nullptr, grid, &h->pcol, h->stream, &okc, h->sw.verbose));
    if (!okr || !okc) { persist_free(h); return KCRN_OK; }
    h->persist_grid = grid;
    KCRN_CUDA(cudaMalloc(&h->gp, sizeof(double) * h->d));
    KCRN_CUDA(cudaMalloc(&h->wpp, sizeof(double) * h->d));
    h->piece_cap = std::max(16, std::max(h->prow.npartial, h->pcol.npartial));
    const size_t slots = (size_t)2 * 2 * grid + (size_t)2 * h->piece_cap;
    KCRN_CUDA(cudaMalloc(&h->pslots, sizeof(uint4) * slots));
    KCRN_CUDA(cudaMemsetAsync(h->pslots, 0, sizeof(uint4) * slots, h->stream));
    KCRN_CUDA(cudaMalloc(&h->pbar, sizeof(unsigned) * 2));
    KCRN_CUDA(cudaMemsetAsync(h->pbar, 0, sizeof(unsigned) * 2, h->stream));
    KCRN_CUDA(cudaStreamSynchronize(h->stream));
    if (h->sw.persist_trace) {
        KCRN_CUDA(cudaMalloc(&h->ptrace, sizeof(long long) * (size_t)grid * kPTraceCap));
        KCRN_CUDA(cudaMemset(h->ptrace, 0, sizeof(long long) * (size_t)grid * kPTraceCap));
    }
    h->persist_ok = true;
    return KCRN_OK;
}

// KCRN_PERSIST_TRACE=1: where the time of the LAST launch went, per interval between the marks of thread 0, mean and
// maximum over the CTAs and the steps (SM clocks; ns at 1.965 GHz)
static void print_trace(kcrn_problem *h) {
    const int G = h->persist_grid;
    std::vector<long long> T((size_t)G * kPTraceCap);
    if (cudaMemcpy(T.data(), h->ptrace, sizeof(long long) * T.size(), cudaMemcpyDeviceToHost) != cudaSuccess) return;
    int nm = 0;
    while (nm < kPTraceCap && T[nm] != 0) ++nm;
    constexpr int NR = 11, NC = 10;
    const char *name[NR + NC] = {"row: wait words", "row: gather, multiply, scatter", "row: basis column", "row: segment sums",
                                 "row: epilogue", "row:   x CTA sums", "row:   x release fence + publish", "row:   x wait for all",
                                 "row:   x acquire fence", "row:   x totals", "row: cut rows, alpha",
                                 "col: wait words", "col: gather, multiply, scatter", "col: segment sums", "col: epilogue",
                                 "col:   x CTA sums", "col:   x release fence + publish", "col:   x wait for all",
                                 "col:   x acquire fence", "col:   x totals", "col: cut columns, beta"};
    double sum[NR + NC] = {0}, mx[NR + NC] = {0}, mn[NR + NC];
    long cnt[NR + NC] = {0};
    for (auto &v : mn) v = 1e30;
    for (int c = 0; c < G; ++c) {
        const long long *t = T.data() + (size_t)c * kPTraceCap;
        int k = 6, pos = 0;            // marks 0 .. 5: start, the four marks of the first exchange, |g|
        while (k < nm) {
            const double d = (double)(t[k] - t[k - 1]);
            sum[pos] += d; cnt[pos]++; mx[pos] = std::max(mx[pos], d); mn[pos] = std::min(mn[pos], d);
            ++k;
            pos = (pos + 1) % (NR + NC);
        }
    }
    fprintf(stderr, "kcrn: persistent kernel trace (last launch, %d marks per CTA; ns at 1.965 GHz)\n", nm);
    double tot = 0;
    for (int i = 0; i < NR + NC; ++i) {
        if (!cnt[i]) continue;
        fprintf(stderr, "kcrn:   %-36s mean %6.0f   min %6.0f   max %6.0f\n", name[i], sum[i] / cnt[i] / 1.965, mn[i] / 1.965, mx[i] / 1.965);
        tot += sum[i] / cnt[i] / 1.965;
    }
    fprintf(stderr, "kcrn:   one row pass + one column pass: %.0f ns\n", tot);
}

int persist_alloc(kcrn_problem *h, int m) {
    if (!h->persist_ok) return KCRN_OK;
    cudaFree(h->Vp);
    h->Vp = nullptr;
    KCRN_CUDA(cudaMalloc(&h->Vp, sizeof(double) * (size_t)m * (size_t)h->d));
    return KCRN_OK;
}

void persist_free(kcrn_problem *h) {
    if (h->ptrace && h->persist_ok) print_trace(h);
    cudaFree(h->ptrace);
    h->ptrace = nullptr;
    free_persist_mat(&h->prow);
    free_persist_mat(&h->pcol);
    cudaFree(h->gp); cudaFree(h->wpp); cudaFree(h->Vp); cudaFree(h->pslots); cudaFree(h->pbar);
    h->gp = h->wpp = h->Vp = nullptr;
    h->pslots = nullptr;
    h->pbar = nullptr;
    h->persist_ok = false;
}

int launch_persist_lanczos(kcrn_problem *h, int m) {
    PersistParams P{};
    P.R = view_of(h->prow, h->csr);
    P.C = view_of(h->pcol, h->csc);
    P.segid = h->csc.segid;
    P.D = h->D; P.t = h->t; P.g = h->g; P.gp = h->gp; P.wp = h->wpp; P.Vp = h->Vp; P.V = h->V;
    P.alphas = h->alphas; P.bet = h->bet; P.ctl = h->ctl; P.info = h->info;
    P.n = h->n; P.d = h->d; P.m = m;
    P.n_samples = (double)h->n_global; P.inv_n = 1.0 / P.n_samples; P.l2 = h->l2;
    P.cta_slots = h->pslots; P.piece_slots = h->pslots + (size_t)2 * 2 * h->persist_grid; P.piece_cap = h->piece_cap;
    P.epoch_base = h->pbar;
    P.trace = h->ptrace;
    P.flags = h->sw.persist_flags;
    // no programmatic launch attribute: all CTAs must become resident, nothing may be waiting for this kernel's
    // trigger on the SMs it needs
    KCRN_CUDA(launch_kernel(persist_lanczos_kernel, dim3(h->persist_grid), dim3(kPThreads), kPSmemBytes, h->stream, false, P));
    h->counters[0] += 1;
    h->counters[3] += m;
    h->counters[4] += m - 1;
    return KCRN_OK;
}

}  // namespace kcrn
