// K5 / K6 / K8 and the small reductions: everything on d- or n-length vectors around the sweeps.
//
// Lanczos (optimizer/cubic.py:77-111) is advanced entirely on the device: a control block in global
// memory (LanczosCtl) carries the column index, the breakdown state and the bookkeeping the reference
// does in Python (:98-99 break, :105-108 truncation, :109 last alpha), so a whole m-step recurrence is
// one CUDA graph with no host round trip.  Every reduction is two-level with a fixed grid and a
// self-resetting arrival ticket: CTA partials are written to global memory and the last CTA to arrive
// adds them in index order -- deterministic, no floating-point atomics.
//
//   update stage 0 : w' = w - alpha v              ; beta = |w'|            (cubic.py:95-97)
//   update stage 1 : w' = w - alpha v              ; c  = V[:, :j+1]' w'    (re-orthogonalisation, K6;
//   update stage 2 : w' = w' - V[:, :j+1] c        ; c2 = V[:, :j+1]' w'     not in the reference,
//   update stage 3 : w' = w' - V[:, :j+1] c2       ; beta = |w'|             SURVEY.md section 8(c))
//   post           : w = hv - beta_prev v_prev     ; alpha = v . w          (cubic.py:93-94, :109)
//                    used after the allreduce of row-sharded runs and by the step-wise API
#include <algorithm>

#include "kcrn_device.cuh"
#include "kcrn_internal.cuh"

namespace kcrn {

namespace {

// Receive buffers of the NVLink peer exchange (kcrn_p2p_attach): every rank's buffer mapped into this process,
// mine, the count of exchange steps done so far (device-side; its parity picks the buffer half), a ticket for
// the kernels that need "last CTA" to bump it.  Slot layout: [half][source rank][d + 2].
struct PeerSet {
    uint4 *recv[kMaxPeers];
    uint4 *mine;
    unsigned *epoch, *ticket;
    int rank, world;
    int64_t stride;
};

struct VecParams {
    LanczosCtl *ctl;
    double *V;
    const double *w_in;
    double *w_out;
    double *wp;
    const double *g;
    int64_t d;
    double *alphas, *bet;
    const double *coef_in;
    double *coef_out;
    double *red_partials;
    unsigned *red_ticket;
    int stage, m, raw;
    int from_wp;               // reorth stage 1 inside the fused loop: w' was already formed by the column sweep
    const double *alpha_raw, *alpha_raw2;   // row-sharded fused loop: sum_i t_i (Av)_i (allreduced) and |v|^2
    double inv_n;
    // fused epilogue + allreduce over NVLink peer memory (kcrn_p2p_attach): every rank's receive buffer mapped
    // into this process, mine, and my step counter
    PeerSet peer;
    double n_samples, l2;
    int32_t *info;
    // row-sharded fused loop with deferred sums (defer_slot >= 0): the alpha pieces the row sweep left (2 x rows_count, +
    // the shares of its long rows), where this kernel leaves its own pieces of |w'|^2
    // d-sharded re-orthogonalisation (row-sharded runs): this rank works on the slice [lo, hi) of the
    // d-vectors only (0, d: everything); sliced != 0: dots and norms are sums over the ranks (finished outside), stage 3
    // leaves zeros outside the slice so that a sum over the ranks re-assembles w'
    int64_t lo, hi;
    int sliced;
    double *slice_scal;        // [2] this rank's share of |w'|^2 (+ pad), summed over the ranks before advance_kernel
    int defer_slot;
    const double *rows_part, *rows_longdot;
    int rows_count, rows_nlong;
    double *post_part;
};

// Sum `v` over the whole grid.  Returns true in every thread of the CTA that arrived last; in that
// CTA thread 0 receives the total.  `sh` holds blockDim.x / 32 doubles.
__device__ __forceinline__ bool grid_sum(double v, double *partials, unsigned *ticket, double *sh, int *s_last,
                                         double *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    v = warp_sum(v);
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double cta = 0.0;
        for (int w = 0; w < nwarp; ++w) cta += sh[w];
        partials[blockIdx.x] = cta;
        __threadfence();
        *s_last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;
    }
    __syncthreads();
    if (!*s_last) return false;
    __threadfence();
    double t = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += __ldcg(partials + i);
    t = warp_sum(t);
    if (lane == 0) sh[warp] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < nwarp; ++w) tot += sh[w];
        *total = tot;
    }
    return true;
}

// |g| into bet[0]; reset of alphas, betas and the control block for an m-step run
__global__ void __launch_bounds__(kVecThreads) lanczos_init_kernel(VecParams P) {
    pdl_wait();
    pdl_trigger();
    __shared__ double sh[kVecThreads / 32];
    __shared__ int s_last;
    double loc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.d; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = P.g[i];
        loc = fma(x, x, loc);
    }
    double tot = 0.0;
    if (!grid_sum(loc, P.red_partials, P.red_ticket, sh, &s_last, &tot)) return;
    for (int i = threadIdx.x; i < P.m; i += blockDim.x) {
        P.alphas[i] = 0.0;
        P.bet[i + 1] = 0.0;
    }
    if (threadIdx.x == 0) {
        P.bet[0] = sqrt(tot);
        LanczosCtl c;
        c.col = 0; c.redo = 0; c.final_ = (P.m == 1); c.done = 0; c.m_eff = P.m; c.nvalid = 1;
        c.store_idx = P.m - 1; c.m = P.m; c.broke = 0; c.hvps = 0; c.last_beta = 0.0;
        c.cur_beta = sqrt(tot); c.cur_alpha = 0.0;
        *P.ctl = c;
    }
}

// V[col] = source / bet[col] outside the fused row sweep (step-wise API)
__global__ void __launch_bounds__(kVecThreads) normalise_kernel(VecParams P) {
    pdl_wait();
    pdl_trigger();
    if (P.ctl->done || P.ctl->redo) return;
    const int col = P.ctl->col;
    const double *src = col == 0 ? P.g : P.wp;
    const double beta = P.ctl->cur_beta, rinv = 1.0 / beta;
    double *vout = P.V + (size_t)col * P.d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.d; i += (int64_t)gridDim.x * blockDim.x)
        vout[i] = quotient(src[i], beta, rinv);
}

template <int E>
__global__ void __launch_bounds__(kVecThreads) lanczos_post_kernel(VecParams P) {
    pdl_wait();
    pdl_trigger();
    __shared__ double sh[kVecThreads / 32];
    __shared__ int s_last;
    if (P.ctl->done) return;
    const int col = P.ctl->col, store_idx = P.ctl->store_idx;
    const bool final_ = P.ctl->final_ != 0;
    const double *v = P.V + (size_t)col * P.d;
    const bool subtract = col > 0 && !final_;
    const double *vpre = subtract ? P.V + (size_t)(col - 1) * P.d : nullptr;
    const double beta_prev = subtract ? P.ctl->cur_beta : 0.0;
    const int64_t base = (int64_t)blockIdx.x * (kVecThreads * E) + threadIdx.x;
    double loc = 0.0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int64_t c = base + (int64_t)e * kVecThreads;
        if (c < P.d) {
            const double vc = v[c];
            double hv = P.w_in[c];
            if (P.raw) hv = __dadd_rn(hv / P.n_samples, __dmul_rn(P.l2, vc));
            if (subtract) hv = __dsub_rn(hv, __dmul_rn(beta_prev, vpre[c]));
            P.w_out[c] = hv;
            loc = __dadd_rn(loc, __dmul_rn(vc, hv));
        }
    }
    double tot = 0.0;
    if (!grid_sum(loc, P.red_partials, P.red_ticket, sh, &s_last, &tot)) return;
    if (threadIdx.x == 0) {
        P.alphas[final_ ? store_idx : col] = tot;
        P.ctl->cur_alpha = tot;
        P.ctl->hvps += 1;
        if (final_) P.ctl->done = 1;
    }
}

// Row-sharded fused loop, after the allreduce of {A't, sum_i t_i (Av)_i}: alpha = v'Hv, w' = y/n + l2 v -
// beta_prev v_prev - alpha v, beta = |w'|, then the bookkeeping of cubic.py:98-109 -- replicated on every rank.
//
// PEER = true: no NCCL call in front of it -- the kernel itself is the allreduce, in the style of NCCL's
// low-latency protocol.  Every rank PUSHES its partial A't into a slot of every peer's receive buffer with 16-byte
// stores {value.lo, step, value.hi, step} (posted NVLink writes: one-way latency, no fence, no separate flag), then
// polls the slots of its own buffer -- local L2 reads -- until both step words match, and adds the ranks' values
// in rank order (the same order on every rank, so the replicas stay bit-identical) while it applies the
// epilogue.  A thread pushes and polls the same elements, so there is no intra-kernel barrier.  The receive
// buffer is double-buffered by the parity of the step counter: a rank can start step s+2 only after it received
// everybody's step s+1 data, i.e. after every peer has finished reading step s.
__device__ __forceinline__ void ll_push(uint4 *slot, double v, unsigned step) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(slot), "r"((unsigned)__double2loint(v)),
                 "r"(step), "r"((unsigned)__double2hiint(v)), "r"(step)
                 : "memory");
}

__device__ __forceinline__ double ll_poll(const uint4 *slot, unsigned step) {
    unsigned lo, f0, hi, f1;
    unsigned long long t0 = 0;
    for (unsigned spin = 0;; ++spin) {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(slot) : "memory");
        if (f0 == step && f1 == step) break;
        if ((spin & 0xffffu) == 0xffffu) {       // a peer that died must not hang the box: give up after ~20 s
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 20000000000ull) __trap();
        }
    }
    return __hiloint2double((int)hi, (int)lo);
}

template <int E, bool PEER>
__global__ void __launch_bounds__(kVecThreads) lanczos_post_sharded_kernel(VecParams P) {
    __shared__ double sh[2 * (kVecThreads / 32)];
    __shared__ double s_alpha[kMaxPeers];
    __shared__ int s_last;
    pdl_wait();
    pdl_trigger();
    const bool deferred = P.defer_slot >= 0;
    const LanczosCtl st = *P.ctl;
    if (st.done) return;
    const int col = st.col, store_idx = st.store_idx;
    const bool final_ = st.final_ != 0;
    const int64_t base = (int64_t)blockIdx.x * (kVecThreads * E) + threadIdx.x;
    // this rank's share of sum_i t_i (Av)_i and |v|^2: left whole by the row sweep's tail, or (deferred) still in per-tile
    // pieces, added here by every CTA in the order of the sweep kernels' consumer-side sums
    double mine_alpha, vsq;
    if (deferred) {
        double x = 0.0, y = 0.0;
        const int ce = (P.rows_count + 1) & ~1;
        for (int i = threadIdx.x; i < P.rows_count; i += kVecThreads) {
            x += __ldcg(P.rows_part + i);
            y += __ldcg(P.rows_part + ce + i);
        }
        for (int i = threadIdx.x; i < P.rows_nlong; i += kVecThreads) x += __ldcg(P.rows_longdot + i);
        x = warp_sum(x);
        y = warp_sum(y);
        if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5] = x; sh[kVecThreads / 32 + (threadIdx.x >> 5)] = y; }
        __syncthreads();
        mine_alpha = vsq = 0.0;
        for (int w = 0; w < kVecThreads / 32; ++w) { mine_alpha += sh[w]; vsq += sh[kVecThreads / 32 + w]; }
        __syncthreads();
    } else {
        mine_alpha = P.alpha_raw[0];
        vsq = P.alpha_raw2[0];
    }
    unsigned step = 0;
    const uint4 *mine = nullptr;
    double s1 = 0.0, ymine[E];
    if constexpr (PEER) {
        const PeerSet &S = P.peer;
        step = *S.epoch + 1u;
        mine = S.mine + (size_t)(step & 1u) * S.world * S.stride;
        const size_t my_slot = ((size_t)(step & 1u) * S.world + S.rank) * S.stride;
        if (blockIdx.x == 0 && threadIdx.x < S.world && threadIdx.x != S.rank)
            ll_push(S.recv[threadIdx.x] + my_slot + P.d, mine_alpha, step);
        if (!final_) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int64_t c = base + (int64_t)e * kVecThreads;
                if (c < P.d) {
                    ymine[e] = P.w_in[c];
                    for (int k = 0; k < S.world; ++k)
                        if (k != S.rank) ll_push(S.recv[k] + my_slot + c, ymine[e], step);
                }
            }
        }
        if (threadIdx.x < S.world)
            s_alpha[threadIdx.x] = threadIdx.x == S.rank ? mine_alpha : ll_poll(mine + threadIdx.x * S.stride + P.d, step);
        __syncthreads();
        for (int k = 0; k < S.world; ++k) s1 += s_alpha[k];
    } else {
        s1 = mine_alpha;
    }
    const double alpha = __dadd_rn(quotient(s1, P.n_samples, P.inv_n), __dmul_rn(P.l2, vsq));
    const double *v = P.V + (size_t)col * P.d;
    const bool subtract = col > 0;
    const double *vpre = subtract ? P.V + (size_t)(col - 1) * P.d : nullptr;
    const double beta_prev = subtract ? st.cur_beta : 0.0;
    double loc = 0.0;
    if (!final_) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int64_t c = base + (int64_t)e * kVecThreads;
            if (c < P.d) {
                const double vc = v[c];
                double y;
                if constexpr (PEER) {
                    y = 0.0;
                    for (int k = 0; k < P.peer.world; ++k)      // rank order: identical bits everywhere
                        y += k == P.peer.rank ? ymine[e] : ll_poll(mine + (size_t)k * P.peer.stride + c, step);
                } else {
                    y = P.w_in[c];
                }
                double hv = __dadd_rn(quotient(y, P.n_samples, P.inv_n), __dmul_rn(P.l2, vc));
                if (subtract) hv = __dsub_rn(hv, __dmul_rn(beta_prev, vpre[c]));
                hv = __dsub_rn(hv, __dmul_rn(alpha, vc));
                P.wp[c] = hv;
                loc = fma(hv, hv, loc);
            }
        }
    }
    if (deferred) {
        // |w'|^2 stays in per-CTA pieces for the next row sweep (it advances the recurrence); nobody waits here.  The
        // exchange counter moves once every CTA has read it: a bare atomic, off every critical path
        loc = warp_sum(loc);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = loc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double cta = 0.0;
            for (int w = 0; w < kVecThreads / 32; ++w) cta += sh[w];
            P.post_part[blockIdx.x] = cta;
            if constexpr (PEER) {
                if (atomicInc(P.peer.ticket, gridDim.x - 1) == gridDim.x - 1) *P.peer.epoch = step;
            }
            if (blockIdx.x == 0) {
                P.alphas[final_ ? store_idx : col] = alpha;
                P.ctl->cur_alpha = alpha;
                P.ctl->hvps = st.hvps + 1;
                if (final_) P.ctl->done = 1;
            }
        }
        return;
    }
    double tot = 0.0;
    if (!grid_sum(loc, P.red_partials, P.red_ticket, sh, &s_last, &tot)) return;
    if (threadIdx.x == 0) {
        if constexpr (PEER) *P.peer.epoch = step;
        P.alphas[final_ ? store_idx : col] = alpha;
        P.ctl->cur_alpha = alpha;
        P.ctl->hvps += 1;
        if (final_) P.ctl->done = 1;
        else if (P.stage == 0) lanczos_advance(P.ctl, P.bet, sqrt(tot));   // stage 1: re-orthogonalisation follows
    }
}

// The per-warp shares of the nv dots of a CTA -> the CTA's shares, stored column by column ([column][CTA]) for
// dots_reduce_kernel.  (Until the last session of round 2 the CTA that arrived last added all shares itself: at
// d = 1.36 M that is 1 324 CTAs x nv values walked by nv threads of ONE CTA -- ~28 us per launch whatever nv is, a quarter
// of the re-orthogonalisation at the news20 shape, profiles/r3d_c3_reorth_launches.csv.  One more small kernel with a CTA
// per column costs ~2 us behind a programmatic edge.)  Called by all threads of the CTA after a barrier behind the last
// write to sdots.
template <int kWarps>
__device__ __forceinline__ void finish_dots(const VecParams &P, int nv, const double (*sdots)[kMaxM]) {
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
        double cta = 0.0;
        for (int wv = 0; wv < kWarps; ++wv) cta += sdots[wv][i];
        P.red_partials[(size_t)i * gridDim.x + blockIdx.x] = cta;
    }
}

// coef_out[i] = sum over the G CTAs' shares of column i = blockIdx.x, in a fixed order (thread-strided, warp butterfly,
// warps in order)
__global__ void __launch_bounds__(kVecThreads) dots_reduce_kernel(VecParams P, int G) {
    pdl_wait();
    pdl_trigger();
    __shared__ double sh[kVecThreads / 32];
    if (P.ctl->done) return;
    const int nv = P.ctl->col + 1, i = blockIdx.x;
    if (i >= nv) return;
    const double *p = P.red_partials + (size_t)i * G;
    double t = 0.0;
    for (int b = threadIdx.x; b < G; b += kVecThreads) t += __ldcg(p + b);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    t = warp_sum(t);
    if (lane == 0) sh[warp] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < kVecThreads / 32; ++w) tot += sh[w];
        P.coef_out[i] = tot;
    }
}

// (Round 2 also measured stage 2 with the basis entries it fetches for w' -= V c kept in shared memory for the coefficients
// V' w' that follow -- "staged through shared memory", bit-identical: 24.3 ms (72 KB per CTA) / 28.3 ms (184 KB) against
// 19.8 ms per 50-step recurrence at the news20 shape, profiles/r2y_c3_reorth_staged.txt.  The second use already comes from
// L2; the shared memory only costs resident CTAs.  Removed.)
template <int E>
__global__ void __launch_bounds__(kVecThreads) lanczos_update_kernel(VecParams P) {
    pdl_wait();
    pdl_trigger();
    __shared__ double sh[kVecThreads / 32];
    __shared__ int s_last;
    __shared__ double sdots[kVecThreads / 32][kMaxM];
    __shared__ double scoef[kMaxM];
    if (P.ctl->done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = P.ctl->col;
    const int nv = col + 1;  // basis columns built so far
    const int64_t base = (int64_t)blockIdx.x * (kVecThreads * E) + threadIdx.x;
    double x[E];
    if (P.stage == 1 && P.from_wp) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int64_t c = base + (int64_t)e * kVecThreads;
            x[e] = (c >= P.lo && c < P.hi) ? P.wp[c] : 0.0;
        }
    } else if (P.stage <= 1) {
        const double alpha = P.ctl->cur_alpha;
        const double *v = P.V + (size_t)col * P.d;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int64_t c = base + (int64_t)e * kVecThreads;
            x[e] = (c >= P.lo && c < P.hi) ? __dsub_rn(P.w_in[c], __dmul_rn(alpha, v[c])) : 0.0;
        }
    } else {
        // w' -= V[:, :nv] c: the coefficients come from shared memory and the basis entries are fetched eight columns
        // at a time, so that eight loads are in flight per thread instead of one dependent load per column
        double s[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int64_t c = base + (int64_t)e * kVecThreads;
            x[e] = (c >= P.lo && c < P.hi) ? P.wp[c] : 0.0;
            s[e] = 0.0;
        }
        for (int i = threadIdx.x; i < nv; i += blockDim.x) scoef[i] = P.coef_in[i];
        __syncthreads();
        for (int i0 = 0; i0 < nv; i0 += 8) {
            double vv[8][E];
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int64_t c = base + (int64_t)e * kVecThreads;
                    vv[u][e] = (i0 + u < nv && (c >= P.lo && c < P.hi)) ? P.V[(size_t)(i0 + u) * P.d + c] : 0.0;
                }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (i0 + u < nv) {
                    const double ci = scoef[i0 + u];
#pragma unroll
                    for (int e = 0; e < E; ++e) s[e] = fma(ci, vv[u][e], s[e]);
                }
        }
#pragma unroll
        for (int e = 0; e < E; ++e) x[e] -= s[e];
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int64_t c = base + (int64_t)e * kVecThreads;
        if ((c >= P.lo && c < P.hi)) P.wp[c] = x[e];
        else if (P.sliced && P.stage == 3 && c < P.d) P.wp[c] = 0.0;      // the owner's value arrives with the sum over the ranks
    }

    if (P.stage == 0 || P.stage == 3) {
        double loc = 0.0;
#pragma unroll
        for (int e = 0; e < E; ++e) loc = fma(x[e], x[e], loc);
        double tot = 0.0;
        if (!grid_sum(loc, P.red_partials, P.red_ticket, sh, &s_last, &tot)) return;
        if (threadIdx.x == 0) {
            if (P.sliced) { P.slice_scal[0] = tot; P.slice_scal[1] = 0.0; }    // the other ranks' shares are still missing
            else lanczos_advance(P.ctl, P.bet, sqrt(tot));
        }
        return;
    }

    // stages 1, 2: dots of the updated vector with every basis column built so far (eight columns per batch)
    for (int i0 = 0; i0 < nv; i0 += 8) {
        double vv[8][E];
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int64_t c = base + (int64_t)e * kVecThreads;
                vv[u][e] = (i0 + u < nv && (c >= P.lo && c < P.hi)) ? P.V[(size_t)(i0 + u) * P.d + c] : 0.0;
            }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (i0 + u < nv) {
                double loc = 0.0;
#pragma unroll
                for (int e = 0; e < E; ++e) loc = fma(vv[u][e], x[e], loc);
                loc = warp_sum(loc);
                if (lane == 0) sdots[warp][i0 + u] = loc;
            }
        }
    }
    __syncthreads();
    finish_dots<kVecThreads / 32>(P, nv, sdots);
}

// Stages 1 and 2 with the basis entries of an element in REGISTERS (K6 where it is bandwidth-bound).  The general kernel
// above walks the columns eight at a time, so stage 2 -- w' -= V c, then c2 = V' w' -- has to fetch every basis entry
// twice: at d = 1.36 M the second fetch comes from HBM again (ncu launch list, profiles/r3c_c3_reorth_launches.csv: 207 us
// for stage 2 against 96 us for stage 3 at 50 columns).  Here a thread owns ONE element at a time and keeps its (up to CAP)
// basis entries in registers: all CAP loads are in flight at once, the update uses them, the dots use them again.  The
// host knows an upper bound of the column count from the position of the step in the captured recurrence and picks CAP.
// The 32 lanes' products of a column are added by recursive halving, eight columns at a time -- lane pairs exchange half
// of their columns per step: 7 + 2 value exchanges per eight columns instead of 40 -- in a fixed tree, and the per-warp
// sums accumulate over the CTA's chunks in shared memory; per-CTA shares and the grid-wide sums as above.  w' gets the same bits as from the
// general kernel (same order of the fused multiply-adds); the coefficients differ in the association of their sums.
constexpr int kRegThreads = 128;

// one halving step: the lane keeps the lower (bit clear) or upper (bit set) half of its N values, adds the partner's
template <int N>
__device__ __forceinline__ void halve(double (&p)[N], int lane, int bit) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int k = 0; k < N / 2; ++k) {
        const double keep = up ? p[k + N / 2] : p[k];
        const double send = up ? p[k] : p[k + N / 2];
        p[k] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
}

template <int CAP, int E>
__global__ void __launch_bounds__(kRegThreads, CAP > 32 ? 3 : 4) lanczos_reorth_reg_kernel(VecParams P) {
    static_assert(CAP % 8 == 0 && CAP >= 8 && CAP <= 64, "column capacity");
    pdl_wait();
    pdl_trigger();
    __shared__ double sdots[kRegThreads / 32][kMaxM];       // only [.][0 .. CAP) is used (the tail's signature)
    __shared__ double scoef[CAP];
    if (P.ctl->done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = P.ctl->col;
    const int nv = col + 1;                                  // <= CAP: the host's bound
    const bool update = P.stage == 2;
    for (int i = threadIdx.x; i < CAP; i += kRegThreads) {
        scoef[i] = (update && i < nv) ? P.coef_in[i] : 0.0;
#pragma unroll
        for (int wv = 0; wv < kRegThreads / 32; ++wv) sdots[wv][i] = 0.0;
    }
    __syncthreads();
    // columns are reduced in groups of eight: three halving steps leave lane l with column 4 b4 + 2 b3 + b2 of the group
    // (b = bits of l), two butterflies finish it; the lanes with b1 = b0 = 0 keep the sums
    const int gcol = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    const double alpha = P.ctl->cur_alpha;
    const double *vcur = P.V + (size_t)col * P.d;
    constexpr int kChunks = kVecThreads * E / kRegThreads;   // the CTA covers what a CTA of the general kernel covers
    for (int q = 0; q < kChunks; ++q) {
        const int64_t c = (int64_t)blockIdx.x * (kVecThreads * E) + (int64_t)q * kRegThreads + threadIdx.x;
        const bool in = c >= P.lo && c < P.hi;
        double vv[CAP];
#pragma unroll
        for (int i = 0; i < CAP; ++i) vv[i] = (in && i < nv) ? __ldg(P.V + (size_t)i * P.d + c) : 0.0;
        double x = 0.0;
        if (in) {
            if (update || P.from_wp) x = P.wp[c];
            else x = __dsub_rn(P.w_in[c], __dmul_rn(alpha, vcur[c]));
        }
        if (update) {
            double sacc = 0.0;
#pragma unroll
            for (int i = 0; i < CAP; ++i) sacc = fma(scoef[i], vv[i], sacc);     // zero coefficients beyond nv: exact no-ops
            x -= sacc;
        }
        if (in && (update || !P.from_wp)) P.wp[c] = x;
#pragma unroll
        for (int i = 0; i < CAP; ++i) vv[i] *= x;
        // sums over the 32 lanes, column by column
#pragma unroll
        for (int g = 0; g < CAP / 8; ++g) {
            double(&p8)[8] = *reinterpret_cast<double(*)[8]>(&vv[8 * g]);
            halve<8>(p8, lane, 16);
            double(&p4)[4] = *reinterpret_cast<double(*)[4]>(&vv[8 * g]);
            halve<4>(p4, lane, 8);
            double(&p2)[2] = *reinterpret_cast<double(*)[2]>(&vv[8 * g]);
            halve<2>(p2, lane, 4);
            double t = vv[8 * g];
            t += __shfl_xor_sync(0xffffffffu, t, 2);
            t += __shfl_xor_sync(0xffffffffu, t, 1);
            if ((lane & 3) == 0) sdots[warp][8 * g + gcol] += t;
        }
    }
    __syncthreads();
    finish_dots<kRegThreads / 32>(P, nv, sdots);
}

// d-sharded re-orthogonalisation: |w'|^2 has been summed over the ranks -- the bookkeeping of cubic.py:98-109
__global__ void advance_kernel(LanczosCtl *ctl, double *bet, const double *bsq) {
    pdl_wait();
    if (threadIdx.x == 0 && !ctl->done) lanczos_advance(ctl, bet, sqrt(bsq[0]));
}

struct FinishParams {
    const LanczosCtl *ctl_in;      // where the last kernel of the recurrence left the state
    LanczosCtl *ctl_main;          // h->ctl[0]: where everybody else looks
    int32_t *info;
    double *bet, *alphas;
    // deferred mode: the alpha of the last operator application is still in per-CTA pieces
    int defer;
    const double *part, *longdot;
    int count, nlong;
    double n_samples, inv_n, l2;
};

__global__ void __launch_bounds__(kVecThreads) lanczos_finish_kernel(FinishParams P) {
    __shared__ double sh[2 * (kVecThreads / 32)];
    pdl_wait();
    pdl_trigger();
    LanczosCtl st = *P.ctl_in;
    if (P.defer && !st.done) {
        // same order as the consumer-side sum of the sweep kernels (thread-strided, warp tree, warps in order)
        double x = 0.0, y = 0.0;
        for (int i = threadIdx.x; i < P.count; i += kVecThreads) {
            x += __ldcg(P.part + i);
            y += __ldcg(P.part + ((P.count + 1) & ~1) + i);
        }
        for (int i = threadIdx.x; i < P.nlong; i += kVecThreads) x += __ldcg(P.longdot + i);
        x = warp_sum(x);
        y = warp_sum(y);
        if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5] = x; sh[kVecThreads / 32 + (threadIdx.x >> 5)] = y; }
        __syncthreads();
        double s1 = 0.0, s2 = 0.0;
        for (int w = 0; w < kVecThreads / 32; ++w) { s1 += sh[w]; s2 += sh[kVecThreads / 32 + w]; }
        const double alpha = __dadd_rn(quotient(s1, P.n_samples, P.inv_n), __dmul_rn(P.l2, s2));
        if (threadIdx.x == 0) P.alphas[st.final_ ? st.store_idx : st.col] = alpha;
        st.cur_alpha = alpha;
        st.hvps += 1;
        if (st.final_) st.done = 1;
    }
    if (threadIdx.x == 0) {
        P.bet[st.m] = st.last_beta;   // slot m is free: bet[j + 1] = beta_j only for j <= m - 2
        P.info[0] = st.m_eff; P.info[1] = st.nvalid; P.info[2] = st.broke; P.info[3] = st.hvps;
        P.info[4] = st.col; P.info[5] = st.done; P.info[6] = st.redo; P.info[7] = st.final_;
        *P.ctl_main = st;
    }
}

// mode 0: sum a[i];  mode 1: sum a[i]^2.   out[0] = total * scale
__global__ void __launch_bounds__(kVecThreads) reduce_kernel(const double *__restrict__ a, int64_t len, int mode,
                                                             double scale, double *partials, unsigned *ticket,
                                                             double *out) {
    __shared__ double sh[kVecThreads / 32];
    __shared__ int s_last;
    double loc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = a[i];
        loc = mode ? fma(x, x, loc) : loc + x;
    }
    double tot = 0.0;
    if (!grid_sum(loc, partials, ticket, sh, &s_last, &tot)) return;
    if (threadIdx.x == 0) out[0] = tot * scale;
}

// value = (sum of loss terms) / n + l2 / 2 * |x|^2   (optimizer/loss.py:215-220)
__global__ void value_combine_kernel(const double *scal, double n_samples, double l2, double *out) {
    if (threadIdx.x == 0) {
        double v = scal[1] / n_samples;
        if (l2 != 0.0) {
            const double nx = sqrt(scal[2]);
            v = __dadd_rn(v, __dmul_rn(l2 / 2.0, __dmul_rn(nx, nx)));
        }
        out[0] = v;
    }
}

// the scalars of one trial, gathered into 12 consecutive doubles so that the host needs ONE copy:
// {value, model decrease, r, Newton iterations, status, m_eff, |g|, operator applications, |x_trial - x|,
//  basis columns written, breakdown flag, 0}
__global__ void record_kernel(const double *__restrict__ scal, const double *__restrict__ sub_tail,
                              const double *__restrict__ bet, const int32_t *__restrict__ info,
                              double *__restrict__ rec) {
    pdl_wait();
    if (threadIdx.x == 0) {
        rec[0] = scal[0]; rec[1] = sub_tail[1]; rec[2] = sub_tail[0]; rec[3] = sub_tail[2]; rec[4] = sub_tail[3];
        rec[5] = (double)info[0]; rec[6] = bet[0]; rec[7] = (double)info[3]; rec[8] = scal[3];
        rec[9] = (double)info[1]; rec[10] = (double)info[2]; rec[11] = 0.0;
    }
}

// value (loss.py:215-220) straight from the forward sweep's per-tile sums -- tiles in tile order, then the shares of the
// long rows: one kernel instead of reduce + combine (+ record)
__global__ void __launch_bounds__(kVecThreads) value_record_kernel(const double *__restrict__ part, int count,
                                                                   const double *__restrict__ longdot, int nlong,
                                                                   const double *__restrict__ scal_in, double n_samples,
                                                                   double l2, double *__restrict__ out,
                                                                   const double *__restrict__ sub_tail,
                                                                   const double *__restrict__ bet,
                                                                   const int32_t *__restrict__ info, double *__restrict__ scal,
                                                                   double *__restrict__ rec) {
    __shared__ double sh[kVecThreads / 32];
    pdl_wait();
    double x = 0.0;
    for (int i = threadIdx.x; i < count; i += kVecThreads) x += __ldcg(part + i);
    for (int i = threadIdx.x; i < nlong; i += kVecThreads) x += __ldcg(longdot + i);
    x = warp_sum(x);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < kVecThreads / 32; ++w) tot += sh[w];
        double v = tot / n_samples;
        if (l2 != 0.0) {
            const double nx = sqrt(scal_in[2]);
            v = __dadd_rn(v, __dmul_rn(l2 / 2.0, __dmul_rn(nx, nx)));
        }
        out[0] = v;
        if (rec) {
            rec[0] = v; rec[1] = sub_tail[1]; rec[2] = sub_tail[0]; rec[3] = sub_tail[2]; rec[4] = sub_tail[3];
            rec[5] = (double)info[0]; rec[6] = bet[0]; rec[7] = (double)info[3]; rec[8] = scal[3];
            rec[9] = (double)info[1]; rec[10] = (double)info[2]; rec[11] = 0.0;
        }
    }
}

// out = y / n + l2 u   (after the allreduce of a row-sharded transpose sweep)
__global__ void scale_kernel(const double *__restrict__ y, const double *__restrict__ u, int64_t d,
                             double n_samples, double l2, double *__restrict__ out) {
    pdl_wait();
    pdl_trigger();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d; i += (int64_t)gridDim.x * blockDim.x) {
        double v = y[i] / n_samples;
        if (u) v = __dadd_rn(v, __dmul_rn(l2, u[i]));
        out[i] = v;
    }
}

// the same with the sum over the ranks inside: y is this rank's partial A' r; push, poll, add in rank order
__global__ void __launch_bounds__(kVecThreads) scale_peer_kernel(PeerSet S, const double *__restrict__ y,
                                                                 const double *__restrict__ u, int64_t d,
                                                                 double n_samples, double l2, double *__restrict__ out) {
    pdl_wait();
    pdl_trigger();
    const unsigned step = *S.epoch + 1u;
    const uint4 *mine = S.mine + (size_t)(step & 1u) * S.world * S.stride;
    const size_t my_slot = ((size_t)(step & 1u) * S.world + S.rank) * S.stride;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, hop = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = first; i < d; i += hop) {
        const double v = y[i];
        for (int k = 0; k < S.world; ++k)
            if (k != S.rank) ll_push(S.recv[k] + my_slot + i, v, step);
    }
    for (int64_t i = first; i < d; i += hop) {
        double tot = 0.0;
        for (int k = 0; k < S.world; ++k) tot += k == S.rank ? y[i] : ll_poll(mine + (size_t)k * S.stride + i, step);
        double v = tot / n_samples;
        if (u) v = __dadd_rn(v, __dmul_rn(l2, u[i]));
        out[i] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0 && atomicInc(S.ticket, gridDim.x - 1) == gridDim.x - 1) *S.epoch = step;
}

// value with the sum of the ranks' loss terms inside (one warp; slot d + 1 of the exchange buffer)
__global__ void value_combine_peer_kernel(PeerSet S, const double *scal, int64_t d, double n_samples, double l2,
                                          double *out) {
    const unsigned step = *S.epoch + 1u;
    const uint4 *mine = S.mine + (size_t)(step & 1u) * S.world * S.stride;
    const size_t my_slot = ((size_t)(step & 1u) * S.world + S.rank) * S.stride;
    const int lane = threadIdx.x;
    double part = 0.0;
    if (lane < S.world) {
        if (lane != S.rank) ll_push(S.recv[lane] + my_slot + d + 1, scal[1], step);
        part = lane == S.rank ? scal[1] : ll_poll(mine + (size_t)lane * S.stride + d + 1, step);
    }
    double tot = 0.0;
    for (int k = 0; k < S.world; ++k) tot += __shfl_sync(0xffffffffu, part, k);
    if (lane == 0) {
        double v = tot / n_samples;
        if (l2 != 0.0) {
            const double nx = sqrt(scal[2]);
            v = __dadd_rn(v, __dmul_rn(l2 / 2.0, __dmul_rn(nx, nx)));
        }
        out[0] = v;
        *S.epoch = step;
    }
}

// out = x + sum_{i < cols} s[i] V[i]     (cubic.py:291,301); optionally |out - x| for the stopping rule
// of the step loop (optimizer.py:110)
__global__ void __launch_bounds__(kVecThreads) axpy_basis_kernel(const double *__restrict__ x,
                                                                 const double *__restrict__ V,
                                                                 const double *__restrict__ s, int m,
                                                                 const int32_t *__restrict__ info, int64_t d,
                                                                 double *__restrict__ out, double *partials,
                                                                 unsigned *ticket, double *step_norm) {
    __shared__ double ss[kMaxM];
    __shared__ double sh[kVecThreads / 32];
    __shared__ int s_last;
    int cols = m;
    if (info) cols = min(info[0], info[1]);
    for (int i = threadIdx.x; i < cols; i += blockDim.x) ss[i] = s[i];
    __syncthreads();
    double loc = 0.0;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < d; c += (int64_t)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int i = 0; i < cols; ++i) acc = fma(V[(size_t)i * d + c], ss[i], acc);
        const double xc = x[c];
        const double xn = xc + acc;
        out[c] = xn;
        const double df = xn - xc;
        loc = fma(df, df, loc);
    }
    if (!step_norm) return;
    double tot = 0.0;
    if (!grid_sum(loc, partials, ticket, sh, &s_last, &tot)) return;
    if (threadIdx.x == 0) step_norm[0] = sqrt(tot);
}

VecParams base_params(kcrn_problem *h) {
    VecParams P{};
    P.ctl = h->ctl; P.V = h->V; P.wp = h->wp; P.g = h->g; P.d = h->d; P.alphas = h->alphas; P.bet = h->bet;
    P.red_partials = h->red_partials; P.n_samples = (double)h->n_global; P.l2 = h->l2; P.info = h->info;
    P.defer_slot = -1;
    P.lo = 0; P.hi = h->d; P.sliced = 0; P.slice_scal = h->scal + 18;
    return P;
}

PeerSet peer_set(kcrn_problem *h) {
    PeerSet S{};
    for (int k = 0; k < h->world; ++k) S.recv[k] = h->peer_recv[k];
    S.mine = h->p2p_recv; S.epoch = h->p2p_epoch; S.ticket = h->red_ticket + 6;
    S.rank = h->rank; S.world = h->world; S.stride = h->d + 2;
    return S;
}

int grid_stride_blocks(int64_t len) {
    int64_t b = (len + kVecThreads - 1) / kVecThreads;
    if (b < 1) b = 1;
    if (b > kNumSm * 8) b = kNumSm * 8;
    return (int)b;
}

}  // namespace

int lanczos_alloc(kcrn_problem *h, int m) {
    if (m <= h->m_cap) return KCRN_OK;
    KCRN_CUDA(cudaStreamSynchronize(h->stream));
    for (auto &ge : h->graphs) cudaGraphExecDestroy(ge.exec);
    h->graphs.clear();
    cudaFree(h->V);
    cudaFree(h->red_partials);
    h->V = nullptr;
    h->red_partials = nullptr;
    KCRN_CUDA(cudaMalloc(&h->V, sizeof(double) * (size_t)m * (size_t)h->d));
    const size_t width = (size_t)m + 1;
    const size_t rows = (size_t)std::max(std::max(h->vec_grid, kNumSm * 8), std::max(h->csc.grid, h->csr.grid));
    KCRN_CUDA(cudaMalloc(&h->red_partials, sizeof(double) * rows * width));
    KCRN_TRY(persist_alloc(h, m));
    h->m_cap = m;
    return KCRN_OK;
}

int launch_lanczos_init(kcrn_problem *h, int m) {
    VecParams P = base_params(h);
    P.m = m;
    P.red_ticket = h->red_ticket + 4;
    KCRN_CUDA(launch_kernel(lanczos_init_kernel, dim3(grid_stride_blocks(h->d)), dim3(kVecThreads), 0, h->stream, h->pdl, P));
    h->counters[0] += 1;
    return KCRN_OK;
}

int launch_normalise_current(kcrn_problem *h) {
    VecParams P = base_params(h);
    normalise_kernel<<<grid_stride_blocks(h->d), kVecThreads, 0, h->stream>>>(P);
    KCRN_LAUNCH_CHECK();
    h->counters[0] += 1;
    return KCRN_OK;
}

int launch_lanczos_post(kcrn_problem *h, const double *hv_or_raw, int raw) {
    VecParams P = base_params(h);
    P.w_in = hv_or_raw; P.w_out = h->w; P.raw = raw;
    P.red_ticket = h->red_ticket + 2;
    if (h->vec_elems == 4) KCRN_CUDA(launch_kernel(lanczos_post_kernel<4>, dim3(h->vec_grid), dim3(kVecThreads), 0, h->stream, h->pdl, P));
    else KCRN_CUDA(launch_kernel(lanczos_post_kernel<1>, dim3(h->vec_grid), dim3(kVecThreads), 0, h->stream, h->pdl, P));
    h->counters[0] += 1;
    return KCRN_OK;
}

// slot >= 0: the fused loop with deferred sums (the control block ping-pongs, see launch_rows_lanczos)
static void post_deferred(kcrn_problem *h, VecParams &P, int slot) {
    P.defer_slot = slot;
    if (slot < 0) return;
    P.ctl = h->ctl + ((slot + 1) & 1);
    P.rows_part = h->lan_rows_part; P.rows_count = h->csr.grid; P.rows_longdot = h->csr.longdot; P.rows_nlong = h->csr.nlong;
    P.post_part = h->post_part;
}

int launch_lanczos_post_sharded(kcrn_problem *h, const double *y_allreduced, int reorth, int slot) {
    VecParams P = base_params(h);
    P.w_in = y_allreduced; P.alpha_raw = h->ypart + h->d; P.alpha_raw2 = h->scal + 17; P.inv_n = 1.0 / P.n_samples;
    P.stage = reorth ? 1 : 0;
    P.red_ticket = h->red_ticket + 2;
    post_deferred(h, P, slot);
    if (h->vec_elems == 4) KCRN_CUDA(launch_kernel(lanczos_post_sharded_kernel<4, false>, dim3(h->vec_grid), dim3(kVecThreads), 0, h->stream, h->pdl, P));
    else KCRN_CUDA(launch_kernel(lanczos_post_sharded_kernel<1, false>, dim3(h->vec_grid), dim3(kVecThreads), 0, h->stream, h->pdl, P));
    h->counters[0] += 1;
    return KCRN_OK;
}

// the same with the allreduce done inside the kernel over peer memory; y_partial = this rank's A't, alpha
// partial behind it at [d]
int launch_lanczos_post_peer(kcrn_problem *h, const double *y_partial, int reorth, int slot) {
    VecParams P = base_params(h);
    P.w_in = y_partial; P.alpha_raw = y_partial + h->d; P.alpha_raw2 = h->scal + 17; P.inv_n = 1.0 / P.n_samples;
    P.stage = reorth ? 1 : 0;
    P.red_ticket = h->red_ticket + 2;
    P.peer = peer_set(h);
    post_deferred(h, P, slot);
    if (h->vec_elems == 4) KCRN_CUDA(launch_kernel(lanczos_post_sharded_kernel<4, true>, dim3(h->vec_grid), dim3(kVecThreads), 0, h->stream, h->pdl, P));
    else KCRN_CUDA(launch_kernel(lanczos_post_sharded_kernel<1, true>, dim3(h->vec_grid), dim3(kVecThreads), 0, h->stream, h->pdl, P));
    h->counters[0] += 1;
    h->counters[5] += 1;
    return KCRN_OK;
}

// reorth = 0: stage 0 (w' = w - alpha v, beta) -- the step-wise API; reorth = 1: the three Gram-Schmidt stages;
// fused = 1: w' already holds w - alpha v (formed by the column sweep / the sharded post kernel)
// max_cols: what the host knows about the number of basis columns at this step (0: nothing)
int launch_lanczos_update(kcrn_problem *h, int reorth, int fused, int max_cols) {
    const int first = reorth ? 1 : 0, last = reorth ? 3 : 0;
    // d-sharded re-orthogonalisation (row-sharded runs with a large basis): every rank orthogonalises its slice of d
    // against the same slice of V -- 1/world of the basis traffic -- and three small sums over the ranks (the
    // coefficients twice, |w'|^2) plus one sum of disjoint slices (w') replace the replicated work
    const bool sliced = reorth && fused && h->dshard_reorth();
    int64_t lo = 0, hi = h->d;
    if (sliced) {
        const int64_t per = (h->d + h->world - 1) / h->world;
        lo = std::min<int64_t>(h->d, per * h->rank);
        hi = std::min<int64_t>(h->d, lo + per);
    }
    for (int stage = first; stage <= last; ++stage) {
        VecParams P = base_params(h);
        P.w_in = h->w; P.stage = stage; P.from_wp = fused;
        P.coef_in = stage == 2 ? h->coef : h->coef2;
        P.coef_out = stage == 1 ? h->coef : h->coef2;
        P.red_ticket = h->red_ticket + 1;
        P.lo = lo; P.hi = hi; P.sliced = sliced ? 1 : 0;
        // stages 1 and 2 of a step with few enough columns: basis entries in registers, every entry fetched once
        // (large d only: at d = 47 k the general kernel is the faster one, 109 vs 116 us per 10-step recurrence --
        // profiles/r3e_reorth_time.txt; KCRN_REORTH_REG=1 forces the register kernel for the tests)
        const bool reg = (stage == 1 || stage == 2) && max_cols > 0 && max_cols <= 64 && !h->sw.no_reorth_reg &&
                         (h->vec_elems == 4 || h->sw.force_reorth_reg);
        auto go = [&](auto kern, int threads) { return launch_kernel(kern, dim3(h->vec_grid), dim3(threads), 0, h->stream, h->pdl, P); };
        if (reg) {
            const int cap = (max_cols + 7) / 8;            // capacity in groups of eight columns
            cudaError_t e = cudaErrorInvalidValue;
#define KCRN_REG_CASE(G)                                                                                              \
    case G: e = h->vec_elems == 4 ? go(lanczos_reorth_reg_kernel<8 * G, 4>, kRegThreads) : go(lanczos_reorth_reg_kernel<8 * G, 1>, kRegThreads); break;
            switch (cap) {
                KCRN_REG_CASE(1) KCRN_REG_CASE(2) KCRN_REG_CASE(3) KCRN_REG_CASE(4)
                KCRN_REG_CASE(5) KCRN_REG_CASE(6) KCRN_REG_CASE(7) KCRN_REG_CASE(8)
                default: break;
            }
#undef KCRN_REG_CASE
            KCRN_CUDA(e);
        } else if (h->vec_elems == 4) KCRN_CUDA(go(lanczos_update_kernel<4>, kVecThreads));
        else KCRN_CUDA(go(lanczos_update_kernel<1>, kVecThreads));
        h->counters[0] += 1;
        if (stage == 1 || stage == 2) {        // the CTAs' shares of the coefficients -> coef_out, one CTA per column
            const int cols = max_cols > 0 ? max_cols : std::max(1, std::max(h->cur_m_enqueue, h->cur_m));
            KCRN_CUDA(launch_kernel(dots_reduce_kernel, dim3(cols), dim3(kVecThreads), 0, h->stream, h->pdl, P, (int)h->vec_grid));
            h->counters[0] += 1;
        }
        if (sliced) {
            if (stage < 3) {
                KCRN_TRY(nccl_allreduce_sum(h, P.coef_out, (h->cur_m_enqueue + 1) & ~1));
            } else {
                KCRN_TRY(nccl_allreduce_sum(h, h->scal + 18, 2));
                KCRN_TRY(nccl_allreduce_sum(h, h->wp, h->d));
                KCRN_CUDA(launch_kernel(advance_kernel, dim3(1), dim3(32), 0, h->stream, h->pdl, h->ctl, h->bet, (const double *)(h->scal + 18)));
                h->counters[0] += 1;
            }
        }
    }
    return KCRN_OK;
}

// m, defer: as the recurrence was enqueued (defer 2: the state ends in h->ctl[m & 1]); defer 0 for the step-wise API
int launch_lanczos_finish(kcrn_problem *h, int m, int defer) {
    static_assert(kVecThreads == kSweepThreads, "the finish kernel repeats the sweep kernels' summation order");
    FinishParams P{};
    P.ctl_in = defer >= 2 ? h->ctl + (m & 1) : h->ctl;     // 2, 3: the state ended in the other half of the ping-pong pair
    if (defer == 3) defer = 0;                             // 3 (row-sharded deferral): the last alpha is already finished
    P.ctl_main = h->ctl; P.info = h->info; P.bet = h->bet; P.alphas = h->alphas;
    P.defer = defer; P.part = h->lan_rows_part; P.longdot = h->csr.longdot; P.count = h->csr.grid; P.nlong = h->csr.nlong;
    P.n_samples = (double)h->n_global; P.inv_n = 1.0 / P.n_samples; P.l2 = h->l2;
    KCRN_CUDA(launch_kernel(lanczos_finish_kernel, dim3(1), dim3(kVecThreads), 0, h->stream, h->pdl, P));
    h->counters[0] += 1;
    return KCRN_OK;
}

int launch_sum(kcrn_problem *h, const double *a, int64_t len, double scale, double *out) {
    reduce_kernel<<<grid_stride_blocks(len), kVecThreads, 0, h->stream>>>(a, len, 0, scale, h->red_partials,
                                                                           h->red_ticket + 3, out);
    KCRN_LAUNCH_CHECK();
    h->counters[0] += 1;
    return KCRN_OK;
}

int launch_sumsq(kcrn_problem *h, const double *a, int64_t len, double *out) {
    reduce_kernel<<<grid_stride_blocks(len), kVecThreads, 0, h->stream>>>(a, len, 1, 1.0, h->red_partials,
                                                                           h->red_ticket + 3, out);
    KCRN_LAUNCH_CHECK();
    h->counters[0] += 1;
    return KCRN_OK;
}

int launch_value_combine(kcrn_problem *h, double *out) {
    value_combine_kernel<<<1, 32, 0, h->stream>>>(h->scal, (double)h->n_global, h->l2, out);
    KCRN_LAUNCH_CHECK();
    h->counters[0] += 1;
    return KCRN_OK;
}

int launch_value_record(kcrn_problem *h, double *out, int with_record) {
    KCRN_CUDA(launch_kernel(value_record_kernel, dim3(1), dim3(kVecThreads), 0, h->stream, h->pdl, (const double *)h->fwd_part,
                            h->csr.grid, (const double *)h->fwd_longdot, h->csr.nlong, (const double *)h->scal,
                            (double)h->n_global, h->l2, out, (const double *)(h->sub + kMaxM), (const double *)h->bet,
                            (const int32_t *)h->info, h->scal, with_record ? h->scal + 4 : (double *)nullptr));
    h->counters[0] += 1;
    return KCRN_OK;
}

int launch_record(kcrn_problem *h) {
    record_kernel<<<1, 32, 0, h->stream>>>(h->scal, h->sub + kMaxM, h->bet, h->info, h->scal + 4);
    KCRN_LAUNCH_CHECK();
    h->counters[0] += 1;
    return KCRN_OK;
}

int launch_scale(kcrn_problem *h, const double *y, const double *u, double *out) {
    KCRN_CUDA(launch_kernel(scale_kernel, dim3(grid_stride_blocks(h->d)), dim3(kVecThreads), 0, h->stream, h->pdl, y, u, h->d, (double)h->n_global, h->l2, out));
    h->counters[0] += 1;
    return KCRN_OK;
}

int launch_scale_peer(kcrn_problem *h, const double *y_partial, const double *u, double *out) {
    KCRN_CUDA(launch_kernel(scale_peer_kernel, dim3(grid_stride_blocks(h->d)), dim3(kVecThreads), 0, h->stream, h->pdl,
                            peer_set(h), y_partial, u, h->d, (double)h->n_global, h->l2, out));
    h->counters[0] += 1;
    h->counters[5] += 1;
    return KCRN_OK;
}

int launch_value_combine_peer(kcrn_problem *h, double *out) {
    value_combine_peer_kernel<<<1, 32, 0, h->stream>>>(peer_set(h), h->scal, h->d, (double)h->n_global, h->l2, out);
    KCRN_LAUNCH_CHECK();
    h->counters[0] += 1;
    h->counters[5] += 1;
    return KCRN_OK;
}

int launch_axpy_basis(const double *x, const double *V, const double *s, int m, const int32_t *info, int64_t d,
                      double *out, cudaStream_t st, double *partials, unsigned *ticket, double *step_norm) {
    axpy_basis_kernel<<<grid_stride_blocks(d), kVecThreads, 0, st>>>(x, V, s, m, info, d, out, partials, ticket,
                                                                     step_norm);
    KCRN_LAUNCH_CHECK();
    return KCRN_OK;
}

}  // namespace kcrn
