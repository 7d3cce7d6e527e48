// f(x_j) for a batch of stored iterates (Trace.compute_loss_of_iterates, optimizer/opt_trace.py:39-43: one
// loss.value(x) -- one sweep of A -- per stored x).  Here the CSR arrays are read ONCE per 8 iterates: the iterates are
// interleaved as XT[c][8], so the gather of column c for all 8 is one 64-byte line; one warp per row, lanes =
// (4 non-zeros) x (8 iterates).  Algorithmic bytes per batch: 12 nnz + 4 (n + 1) + 8 n (labels) + 64 d.
#include "kcrn_internal.cuh"
#include "kcrn_device.cuh"

namespace kcrn {

namespace {

constexpr int kBatch = 8;

// XT[c][j] = X[j0 + j][c] (zero beyond k); sq[j] = |x_j|^2 summed in a fixed order (one CTA per iterate)
__global__ void __launch_bounds__(kVecThreads) interleave_kernel(const double *__restrict__ X, int64_t d, int k, int j0,
                                                                 double *__restrict__ XT, double *__restrict__ sq) {
    __shared__ double sh[kVecThreads / 32];
    const int j = blockIdx.x;
    const bool live = j0 + j < k;
    const double *x = X + (size_t)(j0 + j) * (size_t)d;
    double loc = 0.0;
    for (int64_t c = threadIdx.x; c < d; c += blockDim.x) {
        const double v = live ? x[c] : 0.0;
        XT[c * kBatch + j] = v;
        loc = fma(v, v, loc);
    }
    loc = warp_sum(loc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = loc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < kVecThreads / 32; ++w) tot += sh[w];
        sq[j] = tot;
    }
}

// partial[cta][j] = sum over the CTA's rows of (1 - b_i) z_ij - logsig(z_ij), z_ij = (A x_j)_i; the last CTA adds the
// partials in CTA order: out[j] = total (raw sum; the caller divides by n after the sum over the ranks)
__global__ void __launch_bounds__(kSweepThreads) batch_rows_kernel(const int32_t *__restrict__ ptr,
                                                                   const int32_t *__restrict__ idx,
                                                                   const double *__restrict__ val,
                                                                   const double *__restrict__ b, int64_t n,
                                                                   const double *__restrict__ XT, double *partial,
                                                                   unsigned *ticket, double *__restrict__ out) {
    __shared__ double sh[kSweepWarps][kBatch];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = lane & (kBatch - 1), q = lane >> 3;           // iterate, non-zero slot (4 per pass)
    double acc = 0.0;                                            // lanes 0..7: this warp's sum of loss terms for iterate j
    for (int64_t row = (int64_t)blockIdx.x * kSweepWarps + warp; row < n; row += (int64_t)gridDim.x * kSweepWarps) {
        const int64_t lo = ptr[row], hi = ptr[row + 1];
        double z = 0.0;
        for (int64_t k = lo + q; k < hi; k += 4) z = fma(ld_stream(val + k), XT[(int64_t)ld_stream(idx + k) * kBatch + j], z);
        z += __shfl_xor_sync(0xffffffffu, z, 8);
        z += __shfl_xor_sync(0xffffffffu, z, 16);
        if (q == 0) acc += __dsub_rn(__dmul_rn(1.0 - b[row], z), log_sigmoid(z));
    }
    if (q == 0) sh[warp][j] = acc;
    __syncthreads();
    if (threadIdx.x < kBatch) {
        double tot = 0.0;
        for (int w = 0; w < kSweepWarps; ++w) tot += sh[w][threadIdx.x];
        partial[(size_t)blockIdx.x * kBatch + threadIdx.x] = tot;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < kBatch) {
        double tot = 0.0;
        for (unsigned c = 0; c < gridDim.x; ++c) tot += __ldcg(partial + (size_t)c * kBatch + threadIdx.x);
        out[threadIdx.x] = tot;
    }
}

// value_j = raw_j / n + l2 / 2 |x_j|^2     (optimizer/loss.py:215-220)
__global__ void batch_finish_kernel(const double *raw, const double *sq, int count, double n_samples, double l2, double *out) {
    const int j = threadIdx.x;
    if (j < count) {
        double v = raw[j] / n_samples;
        if (l2 != 0.0) {
            const double nx = sqrt(sq[j]);
            v = __dadd_rn(v, __dmul_rn(l2 / 2.0, __dmul_rn(nx, nx)));
        }
        out[j] = v;
    }
}

}  // namespace

// out[j] = f(X[j]) for j < k; X is [k][d] row-major on the device.  Does not touch the caches of the current x.
int values_batch(kcrn_problem *h, const double *X, int k, double *out) {
    cudaStream_t s = h->stream;
    if (!h->batch_xt) {
        KCRN_CUDA(cudaMalloc(&h->batch_xt, sizeof(double) * kBatch * (size_t)h->d));
        KCRN_CUDA(cudaMalloc(&h->batch_scal, sizeof(double) * 2 * kBatch));
        const int64_t by_rows = (h->n + kSweepWarps - 1) / kSweepWarps;
        h->batch_grid = (int)std::max<int64_t>(1, std::min<int64_t>(by_rows, (int64_t)kNumSm * 8));
        KCRN_CUDA(cudaMalloc(&h->batch_partial, sizeof(double) * kBatch * (size_t)h->batch_grid));
        KCRN_CUDA(cudaMalloc(&h->batch_ticket, sizeof(unsigned)));
        KCRN_CUDA(cudaMemsetAsync(h->batch_ticket, 0, sizeof(unsigned), s));
    }
    for (int j0 = 0; j0 < k; j0 += kBatch) {
        const int count = std::min(kBatch, k - j0);
        interleave_kernel<<<kBatch, kVecThreads, 0, s>>>(X, h->d, k, j0, h->batch_xt, h->batch_scal + kBatch);
        KCRN_LAUNCH_CHECK();
        batch_rows_kernel<<<h->batch_grid, kSweepThreads, 0, s>>>(h->csr.ptr, h->csr.idx, h->csr.val, h->b, h->n, h->batch_xt,
                                                                 h->batch_partial, h->batch_ticket, h->batch_scal);
        KCRN_LAUNCH_CHECK();
        if (h->world > 1) KCRN_TRY(nccl_allreduce_sum(h, h->batch_scal, kBatch));
        batch_finish_kernel<<<1, 32, 0, s>>>(h->batch_scal, h->batch_scal + kBatch, count, (double)h->n_global, h->l2, out + j0);
        KCRN_LAUNCH_CHECK();
        h->counters[0] += 3;
        h->counters[3] += 1;
    }
    return KCRN_OK;
}

}  // namespace kcrn
