// Config-4 rows (full-Hessian CRN and SSCN baselines): the dense contraction A' D A and the column-block
// pieces of SSCN.  Reference: optimizer/loss.py:249-255 (hessian), :234-247 (partial_gradient),
// :257-264 (partial_hessian), :279-281 (update_mat_vec_product).
//
// K9  C = X' Y / n + l2 I  with X = diag(D) A_dense, Y = A_dense, both row-major [K x ld], contracted over the
//     sample index K on the FP64 tensor cores (mma.sync.m8n8k4.f64 -- tcgen05 has no f64 kind).  128 x 128
//     output tile per CTA, 8 warps of 64 x 32, K staged 16 rows at a time through a 3-stage cp.async ring in
//     shared memory (row stride 132 doubles: the four k-lanes of a fragment load land in distinct banks).
//     For the Hessian only tiles on or above the diagonal are computed; each writes its mirror image too.
// K10 column-block gather of A (dense [n x m] block built from the CSC copy), block gradient (one warp per
//     selected column, fixed summation order), and A x += A[:, I] delta (columns applied one after another
//     by a single CTA so the row sums are order-stable without atomics).
#include <algorithm>

#include "kcrn_device.cuh"
#include "kcrn_internal.cuh"

namespace kcrn {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, LDT = BM + 4, STAGES = 3;
constexpr int GEMM_THREADS = 256;
constexpr size_t GEMM_SMEM = (size_t)STAGES * 2 * BK * LDT * sizeof(double);

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma_8x8x4(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// C[p x q] = X' Y / n_samples (+ l2 on the diagonal).  X, Y: [K x ld] row-major, K % 16 == 0, ld % 128 == 0,
// padding zero.  tri != 0: X' Y is symmetric (X = D Y): only tiles with tj >= ti are computed and mirrored.
__global__ void __launch_bounds__(GEMM_THREADS, 1) dgemm_tn_kernel(const double *__restrict__ X,
                                                                  const double *__restrict__ Y, int K, int ld,
                                                                  int p, int q, double n_samples, double l2,
                                                                  double *__restrict__ C, int ldc, int tri,
                                                                  int ksplit, double *__restrict__ part) {
    // ksplit > 1 (few output tiles: SSCN blocks): blockIdx.y = which slice of K this CTA contracts; the raw
    // accumulators go to part[slice][tile][128 x 128] and splitk_reduce_kernel adds the slices in order
    extern __shared__ double smem[];
    int ti, tj;
    if (tri) {
        // linear index over the upper triangle of a T x T tile grid, row by row
        const int T = (p + BM - 1) / BM;
        int rem = blockIdx.x;
        ti = 0;
        while (rem >= T - ti) { rem -= T - ti; ++ti; }
        tj = ti + rem;
    } else {
        ti = blockIdx.y;
        tj = blockIdx.x;
    }
    const int split = tri ? blockIdx.y : 0;
    const int I = ti * BM, J = tj * BN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int gid = lane >> 2, tig = lane & 3;

    auto stage_x = [&](int s) { return smem + (size_t)s * 2 * BK * LDT; };
    auto stage_y = [&](int s) { return smem + (size_t)s * 2 * BK * LDT + BK * LDT; };
    auto load_stage = [&](int s, int k0) {
        double *xs = stage_x(s), *ys = stage_y(s);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = tid + i * GEMM_THREADS;      // 1024 16-byte chunks per operand tile
            const int row = c >> 6, ch = c & 63;
            cp_async_16(xs + row * LDT + ch * 2, X + (size_t)(k0 + row) * ld + I + ch * 2);
            cp_async_16(ys + row * LDT + ch * 2, Y + (size_t)(k0 + row) * ld + J + ch * 2);
        }
    };

    double acc[8][4][2];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    const int nk_all = K / BK, per = (nk_all + ksplit - 1) / ksplit;
    const int kbase = split * per, nk = max(0, min(per, nk_all - kbase));      // this CTA's slice of the K steps
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load_stage(s, (kbase + s) * BK);
        cp_async_commit();
    }
    for (int kc = 0; kc < nk; ++kc) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nxt = kc + STAGES - 1;
        if (nxt < nk) load_stage(nxt % STAGES, (kbase + nxt) * BK);
        cp_async_commit();
        const double *xs = stage_x(kc % STAGES), *ys = stage_y(kc % STAGES);
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double af[8], bf[4];
#pragma unroll
            for (int mt = 0; mt < 8; ++mt) af[mt] = xs[(kk + tig) * LDT + wm * 64 + mt * 8 + gid];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) bf[nt] = ys[(kk + tig) * LDT + wn * 32 + nt * 8 + gid];
#pragma unroll
            for (int mt = 0; mt < 8; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) dmma_8x8x4(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
        }
    }
    cp_async_wait<0>();

#pragma unroll
    for (int mt = 0; mt < 8; ++mt) {
        const int r = I + wm * 64 + mt * 8 + gid;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = J + wn * 32 + nt * 8 + tig * 2 + e;
                if (ksplit > 1) {
                    part[((size_t)split * gridDim.x + blockIdx.x) * (BM * BN) + (size_t)(r - I) * BN + (c - J)] = acc[mt][nt][e];
                    continue;
                }
                if (r < p && c < q) {
                    double v = acc[mt][nt][e] / n_samples;
                    if (r == c && l2 != 0.0) v = __dadd_rn(v, l2);
                    C[(size_t)r * ldc + c] = v;
                    if (tri && ti != tj) C[(size_t)c * ldc + r] = v;   // mirror image of an off-diagonal tile
                }
            }
        }
    }
}

// dense copies of all rows: Y = A, X = diag(w) A.  One thread per row (duplicates add up in order).
__global__ void densify_rows_kernel(const int32_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                                    const double *__restrict__ val, const double *__restrict__ w, int64_t n,
                                    int ld, double *__restrict__ X, double *__restrict__ Y) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const double wr = w[r];
    for (int k = ptr[r]; k < ptr[r + 1]; ++k) {
        const size_t o = (size_t)r * ld + idx[k];
        const double a = val[k];
        Y[o] += a;
        X[o] = __dmul_rn(wr, Y[o]);
    }
}

// dense copy of the selected columns: Y[row, j] = A[row, I[j]], X = diag(w) Y.  One warp per column (a popular column of
// the rcv1-shaped matrix holds 1e5 entries); duplicate entries of one (row, column) add up through the atomic -- the
// sum of two is order-free, canonical matrices have none -- and X is formed once Y is final.
__global__ void densify_cols_kernel(const int32_t *__restrict__ cptr, const int32_t *__restrict__ cidx,
                                    const double *__restrict__ cval, const double *__restrict__ w,
                                    const int64_t *__restrict__ I, int m, int ld, double *__restrict__ X,
                                    double *__restrict__ Y, const int32_t *__restrict__ pos) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (j >= m) return;
    const int64_t c = pos ? pos[I[j]] : I[j];
    const int a = cptr[c], b = cptr[c + 1];
    for (int k = a + lane; k < b; k += 32) atomicAdd(Y + (size_t)cidx[k] * ld + j, cval[k]);
    if (!X) return;
    __threadfence();
    __syncwarp();
    for (int k = a + lane; k < b; k += 32) {
        const size_t o = (size_t)cidx[k] * ld + j;
        X[o] = __dmul_rn(w[cidx[k]], __ldcg(Y + o));
    }
}

// out[j] = sum_k A[k, I[j]] r[k] / n (+ l2 x[I[j]]) -- one warp per selected column
__global__ void partial_gradient_kernel(const int32_t *__restrict__ cptr, const int32_t *__restrict__ cidx,
                                        const double *__restrict__ cval, const double *__restrict__ r,
                                        const int64_t *__restrict__ I, int m, double n_samples, double l2,
                                        const double *__restrict__ x, double *__restrict__ out,
                                        const int32_t *__restrict__ pos) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (j >= m) return;
    const int64_t c = pos ? pos[I[j]] : I[j];
    double acc = 0.0;
    for (int k = cptr[c] + lane; k < cptr[c + 1]; k += 32) acc = fma(cval[k], r[cidx[k]], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
        double g = acc / n_samples;
        if (l2 != 0.0) g = __dadd_rn(g, __dmul_rn(l2, x[I[j]]));
        out[j] = g;
    }
}

// tmp = A[:, I] delta, the columns applied one after another (rows inside a column are distinct)
__global__ void __launch_bounds__(1024) column_update_kernel(const int32_t *__restrict__ cptr,
                                                             const int32_t *__restrict__ cidx,
                                                             const double *__restrict__ cval,
                                                             const int64_t *__restrict__ I,
                                                             const double *__restrict__ delta, int m,
                                                             double *__restrict__ tmp,
                                                             const int32_t *__restrict__ pos) {
    for (int j = 0; j < m; ++j) {
        const int64_t c = pos ? pos[I[j]] : I[j];
        const double dj = delta[j];
        for (int k = cptr[c] + threadIdx.x; k < cptr[c + 1]; k += blockDim.x)
            tmp[cidx[k]] = fma(cval[k], dj, tmp[cidx[k]]);
        __syncthreads();
    }
}

__global__ void add_inplace_kernel(double *__restrict__ a, const double *__restrict__ b, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = a[i] + b[i];
}

// C = (sum over the K slices, in slice order) / n (+ l2 on the diagonal), mirrored below the diagonal
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const double *__restrict__ part, int ksplit, int tiles, int p,
                                                            double n_samples, double l2, double *__restrict__ C, int ldc) {
    const int tile = blockIdx.x, T = (p + BM - 1) / BM;
    int rem = tile, ti = 0;
    while (rem >= T - ti) { rem -= T - ti; ++ti; }
    const int tj = ti + rem;
    for (int e = blockIdx.y * 256 + threadIdx.x; e < BM * BN; e += gridDim.y * 256) {
        const int r = ti * BM + e / BN, c = tj * BN + e % BN;
        if (r >= p || c >= p) continue;
        double acc = 0.0;
        for (int sp = 0; sp < ksplit; ++sp) acc += part[((size_t)sp * tiles + tile) * (BM * BN) + e];
        double v = acc / n_samples;
        if (r == c && l2 != 0.0) v = __dadd_rn(v, l2);
        C[(size_t)r * ldc + c] = v;
        if (ti != tj) C[(size_t)c * ldc + r] = v;
    }
}

int run_gemm(const double *X, const double *Y, int K, int ld, int p, double n_samples, double l2, double *C,
             int ldc, cudaStream_t s, double *part = nullptr, int ksplit = 1) {
    // the attribute is per device: set it every time (a process may drive several GPUs)
    KCRN_CUDA(cudaFuncSetAttribute(dgemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
    const int T = (p + BM - 1) / BM;
    const int64_t tiles = (int64_t)T * (T + 1) / 2;
    dgemm_tn_kernel<<<dim3((unsigned)tiles, (unsigned)ksplit), GEMM_THREADS, GEMM_SMEM, s>>>(X, Y, K, ld, p, p, n_samples, l2, C, ldc,
                                                                                              1, ksplit, part);
    KCRN_LAUNCH_CHECK();
    if (ksplit > 1) {
        splitk_reduce_kernel<<<dim3((unsigned)tiles, 16), 256, 0, s>>>(part, ksplit, (int)tiles, p, n_samples, l2, C, ldc);
        KCRN_LAUNCH_CHECK();
    }
    return KCRN_OK;
}

}  // namespace

int config4_hessian_dense(kcrn_problem *h, double *out) {
    if (!out) { set_error("kcrn_hessian_dense: null output"); return KCRN_E_ARG; }
    if (h->world > 1) { set_error("kcrn_hessian_dense: row-sharded handles are not supported"); return KCRN_E_ARG; }
    const int64_t n = h->n, d = h->d;
    const int ld = (int)((d + BM - 1) / BM * BM);
    const int K = (int)((n + BK - 1) / BK * BK);
    const size_t bytes = sizeof(double) * (size_t)K * (size_t)ld;
    double *X = nullptr, *Y = nullptr;
    KCRN_CUDA(cudaMalloc(&X, bytes));
    if (cudaMalloc(&Y, bytes) != cudaSuccess) {
        cudaFree(X);
        set_error("kcrn_hessian_dense: out of memory for the dense copies (%zu bytes each)", bytes);
        return KCRN_E_CUDA;
    }
    cudaStream_t s = h->stream;
    int rc = KCRN_OK;
    auto run = [&]() -> int {
        KCRN_CUDA(cudaMemsetAsync(X, 0, bytes, s));
        KCRN_CUDA(cudaMemsetAsync(Y, 0, bytes, s));
        densify_rows_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(h->csr.ptr, h->csr.idx, h->csr.val, h->D, n,
                                                                         ld, X, Y);
        KCRN_LAUNCH_CHECK();
        KCRN_TRY(run_gemm(X, Y, K, ld, (int)d, (double)h->n_global, h->l2, out, (int)d, s));
        KCRN_CUDA(cudaStreamSynchronize(s));
        return KCRN_OK;
    };
    rc = run();
    cudaFree(X);
    cudaFree(Y);
    h->counters[0] += 2;
    return rc;
}

int config4_partial_gradient(kcrn_problem *h, const int64_t *I, int m, double *out) {
    if (!I || !out || m < 1) { set_error("kcrn_partial_gradient: bad argument"); return KCRN_E_ARG; }
    if (h->world > 1) { set_error("kcrn_partial_gradient: row-sharded handles are not supported"); return KCRN_E_ARG; }
    partial_gradient_kernel<<<(m * 32 + 255) / 256, 256, 0, h->stream>>>(h->csc.ptr, h->csc.idx, h->csc.val, h->resid,
                                                                        I, m, (double)h->n_global, h->l2, h->x_cur,
                                                                        out, h->csc.pos);
    KCRN_LAUNCH_CHECK();
    h->counters[0] += 1;
    return KCRN_OK;
}

int config4_partial_hessian(kcrn_problem *h, const int64_t *I, int m, double *out) {
    if (!I || !out || m < 1) { set_error("kcrn_partial_hessian: bad argument"); return KCRN_E_ARG; }
    if (h->world > 1) { set_error("kcrn_partial_hessian: row-sharded handles are not supported"); return KCRN_E_ARG; }
    const int64_t n = h->n;
    const int ld = (m + BM - 1) / BM * BM;
    const int K = (int)((n + BK - 1) / BK * BK);
    const size_t bytes = sizeof(double) * (size_t)K * (size_t)ld;
    cudaStream_t s = h->stream;
    // the two dense panels live in the handle (SSCN asks for a block every step): grown on demand, never shrunk
    if (bytes > h->panel_bytes) {
        KCRN_CUDA(cudaStreamSynchronize(s));
        cudaFree(h->panel_x);
        cudaFree(h->panel_y);
        h->panel_x = h->panel_y = nullptr;
        h->panel_bytes = 0;
        KCRN_CUDA(cudaMalloc(&h->panel_x, bytes));
        KCRN_CUDA(cudaMalloc(&h->panel_y, bytes));
        h->panel_bytes = bytes;
    }
    double *X = h->panel_x, *Y = h->panel_y;
    KCRN_CUDA(cudaMemsetAsync(X, 0, bytes, s));
    KCRN_CUDA(cudaMemsetAsync(Y, 0, bytes, s));
    densify_cols_kernel<<<(m * 32 + 127) / 128, 128, 0, s>>>(h->csc.ptr, h->csc.idx, h->csc.val, h->D, I, m, ld, X, Y, h->csc.pos);
    KCRN_LAUNCH_CHECK();
    // a block of m <= 128 columns is ONE output tile: split the contraction over the samples across the SMs (a single
    // CTA walking 20 k rows took 2.5 ms), slices added in order by the reduce kernel
    const int T = (m + BM - 1) / BM, tiles = T * (T + 1) / 2;
    int ksplit = std::max(1, std::min(kNumSm / tiles, (K / BK) / 4));
    if (ksplit > 1) {
        const size_t pbytes = sizeof(double) * (size_t)ksplit * tiles * BM * BN;
        if (pbytes > h->splitk_bytes) {
            KCRN_CUDA(cudaStreamSynchronize(s));
            cudaFree(h->splitk_part);
            h->splitk_part = nullptr;
            h->splitk_bytes = 0;
            KCRN_CUDA(cudaMalloc(&h->splitk_part, pbytes));
            h->splitk_bytes = pbytes;
        }
    }
    KCRN_TRY(run_gemm(X, Y, K, ld, m, (double)h->n_global, h->l2, out, m, s, h->splitk_part, ksplit));
    h->counters[0] += 2;
    return KCRN_OK;
}

int config4_update_Ax(kcrn_problem *h, const int64_t *I, const double *delta, int m) {
    if (!I || !delta || m < 1) { set_error("kcrn_update_Ax: bad argument"); return KCRN_E_ARG; }
    if (h->world > 1) { set_error("kcrn_update_Ax: row-sharded handles are not supported"); return KCRN_E_ARG; }
    cudaStream_t s = h->stream;
    KCRN_CUDA(cudaMemsetAsync(h->t, 0, sizeof(double) * h->n, s));      // t doubles as the A[:, I] delta buffer
    column_update_kernel<<<1, 1024, 0, s>>>(h->csc.ptr, h->csc.idx, h->csc.val, I, delta, m, h->t, h->csc.pos);
    KCRN_LAUNCH_CHECK();
    add_inplace_kernel<<<(unsigned)((h->n + 255) / 256), 256, 0, s>>>(h->Ax, h->t, h->n);
    KCRN_LAUNCH_CHECK();
    h->counters[0] += 2;
    return launch_derive_from_Ax(h);
}

}  // namespace kcrn
