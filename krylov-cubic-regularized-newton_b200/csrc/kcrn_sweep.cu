// K1 / K3 / K4: the sweeps over the sparse data matrix.
//
// One kernel template serves both storage orders.  The matrix is a list of segments (rows of the CSR
// copy, columns of the CSC copy); a host-built, static tile list cuts the non-zeros into pieces of at
// most 256 so that every warp gets the same amount of work whatever the row / column length
// distribution is (power-law columns put a third of the non-zeros of an rcv1-shaped matrix into 100
// columns).  A warp
//   1. loads its (index, value) pairs with fully coalesced loads, 8 in flight per lane,
//   2. gathers the dense vector through the read-only path and multiplies,
//   3. either tree-reduces the 256 products (one piece of a long segment; the piece sums are added in
//      a fixed order by the last piece to arrive -- no floating-point atomics, run-to-run identical),
//      or stages them in shared memory and lets groups of 2^logG lanes sum whole short segments,
//   4. applies the epilogue of the pass in registers (sigmoid weights, 1/n + l2, the Lanczos
//      three-term update and its dot product) so no extra pass over n- or d-vectors is needed.
//
// Reference arithmetic being replaced: optimizer/loss.py:266-277 (A x), :215-232 (value, gradient),
// :289-302 (Hessian-vector product) and the v / w bookkeeping of optimizer/cubic.py:92-97.
#include <algorithm>

#include "kcrn_device.cuh"
#include "kcrn_internal.cuh"

namespace kcrn {

// ----------------------------------------------------------------------------- tile lists (host)
void free_segmat(SegMat *A) {
    cudaFree(A->ptr); cudaFree(A->idx); cudaFree(A->val); cudaFree(A->tiles); cudaFree(A->longtab);
    cudaFree(A->partial); cudaFree(A->longdot); cudaFree(A->ticket);
    *A = SegMat();
}

int build_tiles(SegMat *A, cudaStream_t s) {
    const int64_t nseg = A->nseg;
    std::vector<int32_t> ptr((size_t)nseg + 1);
    KCRN_CUDA(cudaMemcpyAsync(ptr.data(), A->ptr, sizeof(int32_t) * (nseg + 1), cudaMemcpyDeviceToHost, s));
    KCRN_CUDA(cudaStreamSynchronize(s));
    std::vector<int4> tiles, longtab;
    tiles.reserve((size_t)(A->nnz / kTileNnz + nseg / kTileSegs + 16));
    int npartial = 0;
    int64_t sg = 0;
    while (sg < nseg) {
        const int len = ptr[sg + 1] - ptr[sg];
        if (len > kTileNnz) {
            const int pieces = (len + kTileNnz - 1) / kTileNnz;
            const int lid = (int)longtab.size();
            longtab.push_back(make_int4(npartial, pieces, ptr[sg], (int)sg));
            for (int p = 0; p < pieces; ++p) {
                const int k0 = ptr[sg] + p * kTileNnz;
                const int nk = std::min(kTileNnz, ptr[sg + 1] - k0);
                tiles.push_back(make_int4((int)sg, k0, -(lid + 1), nk | (5 << 16)));
            }
            npartial += pieces;
            ++sg;
            continue;
        }
        const int64_t s0 = sg;
        const int k0 = ptr[sg];
        int nk = 0, maxlen = 0;
        while (sg < nseg && sg - s0 < kTileSegs) {
            const int l = ptr[sg + 1] - ptr[sg];
            if (l > kTileNnz || nk + l > kTileNnz) break;
            nk += l;
            maxlen = std::max(maxlen, l);
            ++sg;
        }
        const int ns = (int)(sg - s0);
        int best = 0;
        long best_cost = -1;
        for (int lg = 0; lg <= 5; ++lg) {
            const int G = 1 << lg, groups = 32 >> lg;
            const long rounds = (ns + groups - 1) / groups;
            const long cost = rounds * ((maxlen + G - 1) / G + lg + 2);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = lg; }
        }
        tiles.push_back(make_int4((int)s0, k0, ns, nk | (best << 16)));
    }
    A->ntiles = (int)tiles.size();
    A->nlong = (int)longtab.size();
    A->npartial = npartial;
    A->grid = std::max(1, std::min((A->ntiles + kSweepWarps - 1) / kSweepWarps, kNumSm * 8));
    KCRN_CUDA(cudaMalloc(&A->tiles, sizeof(int4) * std::max<size_t>(tiles.size(), 1)));
    KCRN_CUDA(cudaMalloc(&A->longtab, sizeof(int4) * std::max<size_t>(longtab.size(), 1)));
    KCRN_CUDA(cudaMalloc(&A->partial, sizeof(double) * std::max(npartial, 1)));
    KCRN_CUDA(cudaMalloc(&A->longdot, sizeof(double) * std::max(A->nlong, 1)));
    KCRN_CUDA(cudaMalloc(&A->ticket, sizeof(unsigned) * std::max(A->nlong, 1)));
    KCRN_CUDA(cudaMemsetAsync(A->ticket, 0, sizeof(unsigned) * std::max(A->nlong, 1), s));
    KCRN_CUDA(cudaMemsetAsync(A->longdot, 0, sizeof(double) * std::max(A->nlong, 1), s));
    if (!tiles.empty())
        KCRN_CUDA(cudaMemcpyAsync(A->tiles, tiles.data(), sizeof(int4) * tiles.size(), cudaMemcpyHostToDevice, s));
    if (!longtab.empty())
        KCRN_CUDA(cudaMemcpyAsync(A->longtab, longtab.data(), sizeof(int4) * longtab.size(), cudaMemcpyHostToDevice, s));
    KCRN_CUDA(cudaStreamSynchronize(s));
    return KCRN_OK;
}

// ----------------------------------------------------------------------------- the sweep kernel
namespace {

enum SweepMode {
    kRowsForward = 0,   // A x, sigma, weights, residual, loss terms      (loss.py:215-232, 266-277, 296-297)
    kRowsWeighted = 1,  // t = D .* (A v)                                 (loss.py:299-301)
    kRowsLanczos = 2,   // same; v = source / beta taken from the control block, V[col] written on the side
    kColsRaw = 3,       // out = A' src                                   (row-sharded runs, before the allreduce)
    kColsScaled = 4,    // out = A' src / n (+ l2 u)                      (loss.py:227-229, :302)
    kColsLanczos = 5,   // w = A' t / n + l2 v - beta v_prev ; alpha = v . w   (loss.py:302 + cubic.py:93-94,109)
};

struct SweepParams {
    const int32_t *ptr, *idx;
    const double *val;
    const int4 *tiles;
    int ntiles;
    const int4 *longtab;
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
    double n_samples, l2;
    int add_l2;
    // Lanczos
    LanczosCtl *ctl;
    double *V;
    const double *g, *wp, *bet;
    double *w, *alphas;
    int64_t d;
    double *red_partials;
    unsigned *red_ticket;
};

// Epilogue of one finished segment.  op0 / op1 are the pass's per-segment operands, already in registers:
//   rows forward: label b            rows weighted: Hessian weight D
//   cols scaled : u (l2 term)        cols Lanczos : v[col], v[col-1]
template <int MODE>
__device__ __forceinline__ void finish_segment(const SweepParams &P, int seg, double sum, int lseg, double op0,
                                               double op1, double beta_prev, bool subtract, double &acc) {
    if constexpr (MODE == kRowsForward) {
        const double a = sigmoid(sum);
        P.Ax[seg] = sum;
        P.D[seg] = __dmul_rn(a, 1.0 - a);
        P.resid[seg] = a - op0;
        P.ell[seg] = __dsub_rn(__dmul_rn(1.0 - op0, sum), log_sigmoid(sum));
    } else if constexpr (MODE == kRowsWeighted || MODE == kRowsLanczos) {
        P.out[seg] = __dmul_rn(op0, sum);
    } else if constexpr (MODE == kColsRaw) {
        P.out[seg] = sum;
    } else if constexpr (MODE == kColsScaled) {
        double y = sum / P.n_samples;
        if (P.add_l2) y = __dadd_rn(y, __dmul_rn(P.l2, op0));
        P.out[seg] = y;
    } else {  // kColsLanczos
        double hv = __dadd_rn(sum / P.n_samples, __dmul_rn(P.l2, op0));
        if (subtract) hv = __dsub_rn(hv, __dmul_rn(beta_prev, op1));
        P.w[seg] = hv;
        const double prod = __dmul_rn(op0, hv);
        if (lseg >= 0) P.longdot[lseg] = prod;  // long segments finish on whichever warp arrives last:
        else acc += prod;                       // keep their share out of the per-warp running sums
    }
}

template <int MODE>
__device__ __forceinline__ void load_operands(const SweepParams &P, int seg, const double *vcol, const double *vpre,
                                              bool subtract, double &op0, double &op1) {
    op0 = 0.0;
    op1 = 0.0;
    if constexpr (MODE == kRowsForward) op0 = __ldg(P.b + seg);
    if constexpr (MODE == kRowsWeighted || MODE == kRowsLanczos) op0 = __ldg(P.D + seg);
    if constexpr (MODE == kColsScaled) { if (P.add_l2) op0 = __ldg(P.u + seg); }
    if constexpr (MODE == kColsLanczos) {
        op0 = __ldg(vcol + seg);
        if (subtract) op1 = __ldg(vpre + seg);
    }
}

__host__ __device__ constexpr int kOperands(int mode) { return mode == kColsRaw ? 0 : (mode == kColsLanczos ? 2 : 1); }

template <int MODE>
__global__ void __launch_bounds__(kSweepThreads, 4) sweep_kernel(const SweepParams P) {
    __shared__ double sprod[kSweepWarps][kTileNnz];       // products of the tile
    __shared__ double ssum[kSweepWarps][kTileSegs];       // per-segment sums of the tile
    __shared__ int sptr[kSweepWarps][kTileSegs + 8];      // segment pointers of the tile
    __shared__ double sred[kSweepWarps];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    const double *src = P.src;
    double beta = 1.0, rinv = 1.0;
    const double *vcol = nullptr, *vpre = nullptr;
    double beta_prev = 0.0;
    bool subtract = false, final_ = false;
    int col = 0, store_idx = 0;
    if constexpr (MODE == kRowsLanczos) {
        if (P.ctl->done) return;
        col = P.ctl->col;
        if (P.ctl->redo) {
            src = P.V + (size_t)col * P.d;
        } else {
            src = col == 0 ? P.g : P.wp;
            beta = P.bet[col];
            rinv = 1.0 / beta;
        }
    }
    if constexpr (MODE == kColsLanczos) {
        if (P.ctl->done) return;
        col = P.ctl->col;
        final_ = P.ctl->final_ != 0;
        store_idx = P.ctl->store_idx;
        vcol = P.V + (size_t)col * P.d;
        if (col > 0 && !final_) {
            subtract = true;
            vpre = P.V + (size_t)(col - 1) * P.d;
            beta_prev = P.bet[col];
        }
    }

    double acc = 0.0;
    double *sp = sprod[warp], *sm = ssum[warp];
    int *sq = sptr[warp];
    const int gw = blockIdx.x * kSweepWarps + warp, tw = gridDim.x * kSweepWarps;
    for (int tile = gw; tile < P.ntiles; tile += tw) {
        const int4 td = __ldg(P.tiles + tile);
        const int seg0 = td.x, k0 = td.y, z = td.z, nk = td.w & 0xffff, logG = td.w >> 16;
        const int32_t *ip = P.idx + k0;
        const double *vp = P.val + k0;
        // everything that only depends on the descriptor is requested before anything is consumed
        int ci[8];
        double av[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int i = r * 32 + lane;
            ci[r] = i < nk ? ld_stream(ip + i) : -1;
            av[r] = i < nk ? ld_stream(vp + i) : 0.0;
        }
        int4 lt = make_int4(0, 1, 0, 0);
        constexpr int SR = kTileSegs / 32;   // segment rounds per lane
        if (z < 0) {
            lt = __ldg(P.longtab + (-z - 1));
        } else {
            // segment pointers go straight to shared memory (cp.async: no registers held while in flight)
#pragma unroll
            for (int r = 0; r < SR + 1; ++r) {
                const int i = r * 32 + lane;
                if (i <= z) cp_async_4(sq + i, P.ptr + seg0 + i);
            }
        }
        double p[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            double xv = 0.0;
            if (ci[r] >= 0) xv = __ldg(src + ci[r]);
            if constexpr (MODE == kRowsLanczos) xv = quotient(xv, beta, rinv);
            p[r] = av[r] * xv;
        }
        if (z < 0) {
            // one 256-entry piece of a long segment: tree-reduce, publish, the last piece to arrive finishes
            double op0, op1;
            load_operands<MODE>(P, seg0, vcol, vpre, subtract, op0, op1);
            double s = ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
            s = warp_sum(s);
            const int lseg = -z - 1;
            const int slot = lt.x + ((k0 - lt.z) >> 8);
            int last = 0;
            if (lane == 0) {
                P.partial[slot] = s;
                __threadfence();
                last = atomicInc(P.ticket + lseg, (unsigned)(lt.y - 1)) == (unsigned)(lt.y - 1);
            }
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) {
                __threadfence();
                double tot = 0.0;
                for (int i = lane; i < lt.y; i += 32) tot += __ldcg(P.partial + lt.x + i);
                tot = warp_sum(tot);
                if (lane == 0) finish_segment<MODE>(P, seg0, tot, lseg, op0, op1, beta_prev, subtract, acc);
            }
        } else {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int i = r * 32 + lane;
                if (i < nk) sp[i] = p[r];
            }
            // operands of the epilogue: requested now, consumed after the shared-memory reduction
            constexpr int NOP = kOperands(MODE);
            double o0[NOP > 0 ? SR : 1], o1[NOP > 1 ? SR : 1];
#pragma unroll
            for (int r = 0; r < SR; ++r) {
                const int sg = r * 32 + lane;
                if (sg < z) {
                    double a0, a1;
                    load_operands<MODE>(P, seg0 + sg, vcol, vpre, subtract, a0, a1);
                    if constexpr (NOP > 0) o0[r] = a0;
                    if constexpr (NOP > 1) o1[r] = a1;
                }
            }
            cp_async_wait_all();
            __syncwarp();
            // phase A: every group of G lanes sums one segment from shared memory
            const int G = 1 << logG, groups = 32 >> logG, gid = lane >> logG, gl = lane & (G - 1);
            for (int sb = 0; sb < z; sb += groups) {
                const int sg = sb + gid;
                int a = 0, b = 0;
                if (sg < z) { a = sq[sg] - k0; b = sq[sg + 1] - k0; }
                double sum = 0.0;
                for (int i = a + gl; i < b; i += G) sum += sp[i];
                for (int o = G >> 1; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                if (sg < z && gl == 0) sm[sg] = sum;
            }
            __syncwarp();
            // phase B: coalesced epilogue over the segments of the tile
#pragma unroll
            for (int r = 0; r < SR; ++r) {
                const int sg = r * 32 + lane;
                if (sg < z)
                    finish_segment<MODE>(P, seg0 + sg, sm[sg], -1, NOP > 0 ? o0[r] : 0.0, NOP > 1 ? o1[r] : 0.0,
                                         beta_prev, subtract, acc);
            }
            __syncwarp();
        }
    }

    if constexpr (MODE == kRowsLanczos) {
        // side job: store the normalised vector as basis column `col` (the sweep above divided on the fly)
        if (!P.ctl->redo) {
            double *vout = P.V + (size_t)col * P.d;
            for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < P.d;
                 c += (int64_t)gridDim.x * blockDim.x)
                vout[c] = quotient(__ldg(src + c), beta, rinv);
        }
    }

    if constexpr (MODE == kColsLanczos) {
        // alpha = v . w : lanes -> warp -> CTA -> (last CTA) whole grid, every level in a fixed order
        acc = warp_sum(acc);
        if (lane == 0) sred[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double cta = 0.0;
            for (int wv = 0; wv < kSweepWarps; ++wv) cta += sred[wv];
            P.red_partials[blockIdx.x] = cta;
            __threadfence();
            s_last = atomicInc(P.red_ticket, gridDim.x - 1) == gridDim.x - 1;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            double tot = 0.0;
            for (int i = threadIdx.x; i < (int)gridDim.x; i += kSweepThreads) tot += __ldcg(P.red_partials + i);
            for (int i = threadIdx.x; i < P.nlong; i += kSweepThreads) tot += __ldcg(P.longdot + i);
            tot = warp_sum(tot);
            __syncthreads();
            if (lane == 0) sred[warp] = tot;
            __syncthreads();
            if (threadIdx.x == 0) {
                double alpha = 0.0;
                for (int wv = 0; wv < kSweepWarps; ++wv) alpha += sred[wv];
                P.alphas[final_ ? store_idx : col] = alpha;
                P.ctl->hvps += 1;
                if (final_) P.ctl->done = 1;
            }
        }
    }
}

void fill_matrix(SweepParams &P, const SegMat &A) {
    P.ptr = A.ptr; P.idx = A.idx; P.val = A.val; P.tiles = A.tiles; P.ntiles = A.ntiles;
    P.longtab = A.longtab; P.partial = A.partial; P.longdot = A.longdot; P.ticket = A.ticket;
    P.nlong = A.nlong;
}

template <int MODE>
int launch(kcrn_problem *h, const SegMat &A, SweepParams &P) {
    fill_matrix(P, A);
    sweep_kernel<MODE><<<A.grid, kSweepThreads, 0, h->stream>>>(P);
    KCRN_LAUNCH_CHECK();
    h->counters[0] += 1;
    return KCRN_OK;
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
    h->counters[3] += 1;
    return launch<kRowsForward>(h, h->csr, P);
}

int launch_rows_weighted(kcrn_problem *h, const double *v) {
    SweepParams P{};
    P.src = v; P.D = h->D; P.out = h->t;
    h->counters[3] += 1;
    return launch<kRowsWeighted>(h, h->csr, P);
}

int launch_rows_lanczos(kcrn_problem *h) {
    SweepParams P{};
    P.D = h->D; P.out = h->t; P.ctl = h->ctl; P.V = h->V; P.g = h->g; P.wp = h->wp; P.bet = h->bet; P.d = h->d;
    h->counters[3] += 1;
    return launch<kRowsLanczos>(h, h->csr, P);
}

int launch_cols_raw(kcrn_problem *h, const double *src, double *out) {
    SweepParams P{};
    P.src = src; P.out = out;
    h->counters[4] += 1;
    return launch<kColsRaw>(h, h->csc, P);
}

int launch_cols_scaled(kcrn_problem *h, const double *src, const double *u, double *out) {
    SweepParams P{};
    P.src = src; P.out = out; P.u = u; P.n_samples = (double)h->n_global; P.l2 = h->l2;
    P.add_l2 = u != nullptr;
    h->counters[4] += 1;
    return launch<kColsScaled>(h, h->csc, P);
}

int launch_cols_lanczos(kcrn_problem *h) {
    SweepParams P{};
    P.src = h->t; P.ctl = h->ctl; P.V = h->V; P.bet = h->bet; P.w = h->w; P.alphas = h->alphas; P.d = h->d;
    P.n_samples = (double)h->n_global; P.l2 = h->l2;
    P.red_partials = h->red_partials; P.red_ticket = h->red_ticket + 0;
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
