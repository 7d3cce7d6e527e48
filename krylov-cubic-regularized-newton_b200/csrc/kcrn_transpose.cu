// K2: CSR -> CSC on the device, bit-identical to scipy's csr_tocsc (what the reference gets from
// A.tocsc(), cubic_newton.py:55, and implicitly from A.T in optimizer/loss.py:227,302).
//
// csr_tocsc is a stable counting sort by column: inside a column, entries keep CSR order (row, then
// position in the row); duplicates are kept.  Here the same permutation is produced by a stable
// least-significant-digit radix sort of the pairs (column, CSR position) keyed on the column only:
//   per 8-bit digit:  block histograms -> exclusive scan (digit-major) -> stable scatter
// Ranks inside a block come from warp match/ballot, so no floating-point or order-dependent atomics
// are involved and the result does not depend on scheduling.  The column pointer is read off the
// sorted keys by binary search; row indices and values are gathered through the sorted positions.
#include "kcrn_internal.cuh"

namespace kcrn {

namespace {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// sum of every tile -> sums[block]
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(const int32_t *__restrict__ in,
                                                               int32_t *__restrict__ sums, int64_t len) {
    __shared__ int sh[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    int s = 0;
    for (int i = threadIdx.x; i < kScanTile; i += kScanThreads) {
        const int64_t p = base + i;
        if (p < len) s += in[p];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) tot += sh[w];
        sums[blockIdx.x] = tot;
    }
}

// in-place exclusive scan of a short array by one CTA (1024 threads, chunks with a running carry)
__global__ void __launch_bounds__(1024) scan_single_cta(int32_t *a, int64_t len) {
    __shared__ int wsum[32];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < len; base += 1024) {
        const int64_t p = base + threadIdx.x;
        const int v = p < len ? a[p] : 0;
        int inc = warp_incl_scan(v, lane);
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int ws = wsum[lane];
            int wi = warp_incl_scan(ws, lane);
            wsum[lane] = wi - ws;  // exclusive prefix of the warp totals
        }
        __syncthreads();
        const int excl = carry + wsum[warp] + inc - v;
        if (p < len) a[p] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
}

// exclusive scan inside every tile, shifted by the scanned tile sums
__global__ void __launch_bounds__(kScanThreads) scan_apply(const int32_t *in, int32_t *out,
                                                           const int32_t *__restrict__ offs, int64_t len) {
    __shared__ int wsum[kScanThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int v[kScanItems];
    int s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = (base + i < len) ? in[base + i] : 0;
        s += v[i];
    }
    const int inc = warp_incl_scan(s, lane);
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int before = offs[blockIdx.x];
    for (int w = 0; w < warp; ++w) before += wsum[w];
    int run = before + inc - s;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < len) out[base + i] = run;
        run += v[i];
    }
}

// ---------------------------------------------------------------- radix passes
constexpr int kSortWarps = 8;
constexpr int kSortRounds = 32;
constexpr int kSortTile = kSortWarps * kSortRounds * 32;  // 8192 keys per CTA
constexpr int kSortThreads = kSortWarps * 32;

__global__ void __launch_bounds__(kSortThreads) radix_hist(const int32_t *__restrict__ keys, int64_t n,
                                                           int shift, int32_t *__restrict__ hist, int nblocks) {
    __shared__ int sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
    for (int i = threadIdx.x; i < kSortTile; i += kSortThreads) {
        const int64_t p = base + i;
        if (p < n) atomicAdd(&sh[(keys[p] >> shift) & 255], 1);  // integer counts: order-free
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = sh[threadIdx.x];
}

// Stable scatter.  Order inside the CTA is (warp, round, lane) = ascending input position.
__global__ void __launch_bounds__(kSortThreads) radix_scatter(const int32_t *__restrict__ keys_in,
                                                              const int32_t *__restrict__ pay_in, int64_t n,
                                                              int shift, const int32_t *__restrict__ offs,
                                                              int nblocks, int32_t *__restrict__ keys_out,
                                                              int32_t *__restrict__ pay_out) {
    __shared__ int whist[kSortWarps][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&whist[0][0])[i] = 0;
    __syncthreads();
    const int64_t wbase = (int64_t)blockIdx.x * kSortTile + (int64_t)warp * (kSortRounds * 32);
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int r = 0; r < kSortRounds; ++r) {
        const int64_t p = wbase + r * 32 + lane;
        const bool valid = p < n;
        const unsigned act = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const int dg = (keys_in[p] >> shift) & 255;
            const unsigned peers = __match_any_sync(act, dg);
            if ((peers & lt_mask) == 0) whist[warp][dg] += __popc(peers);
        }
        __syncwarp();
    }
    __syncthreads();
    {
        const int dg = threadIdx.x;  // 256 threads = 256 digits
        int run = offs[(int64_t)dg * nblocks + blockIdx.x];
        for (int w = 0; w < kSortWarps; ++w) {
            const int c = whist[w][dg];
            whist[w][dg] = run;
            run += c;
        }
    }
    __syncthreads();
    for (int r = 0; r < kSortRounds; ++r) {
        const int64_t p = wbase + r * 32 + lane;
        const bool valid = p < n;
        const unsigned act = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const int key = keys_in[p];
            const int pay = pay_in ? pay_in[p] : (int)p;
            const int dg = (key >> shift) & 255;
            const unsigned peers = __match_any_sync(act, dg);
            const int pos = whist[warp][dg] + __popc(peers & lt_mask);
            keys_out[pos] = key;
            pay_out[pos] = pay;
            __syncwarp(peers);
            if ((peers & lt_mask) == 0) whist[warp][dg] += __popc(peers);
        }
        __syncwarp();
    }
}

// column pointer from the sorted column keys: ptr[c] = first slot whose key >= c
__global__ void colptr_from_sorted(const int32_t *__restrict__ keys, int64_t nnz, int64_t n_cols,
                                   int32_t *__restrict__ ptr) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_cols) return;
    int64_t lo = 0, hi = nnz;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < (int32_t)c) lo = mid + 1; else hi = mid;
    }
    ptr[c] = (int32_t)lo;
}

// slot -> (row of the CSR position, value at the CSR position)
__global__ void gather_rows_vals(const int32_t *__restrict__ pos, int64_t nnz,
                                 const int32_t *__restrict__ rowptr, int64_t n_rows,
                                 const double *__restrict__ val, int32_t *__restrict__ out_row,
                                 double *__restrict__ out_val) {
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= nnz) return;
    const int32_t k = pos[slot];
    // largest r with rowptr[r] <= k  (skips empty rows)
    int64_t lo = 0, hi = n_rows;
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if (rowptr[mid] <= k) lo = mid; else hi = mid - 1;
    }
    out_row[slot] = (int32_t)lo;
    out_val[slot] = val[k];
}

}  // namespace

int exclusive_scan_i32(const int32_t *in, int32_t *out, int64_t len, cudaStream_t s) {
    if (len <= 0) return KCRN_OK;
    const int64_t nblk = (len + kScanTile - 1) / kScanTile;
    int32_t *sums = nullptr;
    KCRN_CUDA(cudaMallocAsync(&sums, sizeof(int32_t) * nblk, s));
    scan_tile_sums<<<(unsigned)nblk, kScanThreads, 0, s>>>(in, sums, len);
    KCRN_LAUNCH_CHECK();
    scan_single_cta<<<1, 1024, 0, s>>>(sums, nblk);
    KCRN_LAUNCH_CHECK();
    scan_apply<<<(unsigned)nblk, kScanThreads, 0, s>>>(in, out, sums, len);
    KCRN_LAUNCH_CHECK();
    KCRN_CUDA(cudaFreeAsync(sums, s));
    return KCRN_OK;
}

int build_csc(const SegMat &csr, int64_t n_cols, SegMat *csc, cudaStream_t s) {
    const int64_t nnz = csr.nnz, n_rows = csr.nseg;
    csc->nseg = n_cols;
    csc->nnz = nnz;
    KCRN_CUDA(cudaMalloc(&csc->ptr, sizeof(int32_t) * (n_cols + 1)));
    KCRN_CUDA(cudaMalloc(&csc->idx, sizeof(int32_t) * (nnz > 0 ? nnz : 1)));
    KCRN_CUDA(cudaMalloc(&csc->val, sizeof(double) * (nnz > 0 ? nnz : 1)));
    if (nnz == 0) {
        KCRN_CUDA(cudaMemsetAsync(csc->ptr, 0, sizeof(int32_t) * (n_cols + 1), s));
        return KCRN_OK;
    }
    int bits = 1;
    while (bits < 31 && ((int64_t)1 << bits) < n_cols) ++bits;
    const int passes = (bits + 7) / 8;
    const int nblocks = (int)((nnz + kSortTile - 1) / kSortTile);

    int32_t *keys[2] = {nullptr, nullptr}, *pay[2] = {nullptr, nullptr}, *hist = nullptr;
    for (int i = 0; i < 2; ++i) {
        KCRN_CUDA(cudaMalloc(&keys[i], sizeof(int32_t) * nnz));
        KCRN_CUDA(cudaMalloc(&pay[i], sizeof(int32_t) * nnz));
    }
    const int64_t hlen = (int64_t)256 * nblocks;
    KCRN_CUDA(cudaMalloc(&hist, sizeof(int32_t) * hlen));

    const int32_t *kin = csr.idx;
    const int32_t *pin = nullptr;  // first pass: payload = CSR position itself
    int cur = 0;
    for (int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
        radix_hist<<<nblocks, kSortThreads, 0, s>>>(kin, nnz, shift, hist, nblocks);
        KCRN_LAUNCH_CHECK();
        KCRN_TRY(exclusive_scan_i32(hist, hist, hlen, s));
        radix_scatter<<<nblocks, kSortThreads, 0, s>>>(kin, pin, nnz, shift, hist, nblocks, keys[cur], pay[cur]);
        KCRN_LAUNCH_CHECK();
        kin = keys[cur];
        pin = pay[cur];
        cur ^= 1;
    }
    colptr_from_sorted<<<(unsigned)((n_cols + 1 + 255) / 256), 256, 0, s>>>(kin, nnz, n_cols, csc->ptr);
    KCRN_LAUNCH_CHECK();
    gather_rows_vals<<<(unsigned)((nnz + 255) / 256), 256, 0, s>>>(pin, nnz, csr.ptr, n_rows, csr.val,
                                                                     csc->idx, csc->val);
    KCRN_LAUNCH_CHECK();
    KCRN_CUDA(cudaStreamSynchronize(s));
    for (int i = 0; i < 2; ++i) {
        cudaFree(keys[i]);
        cudaFree(pay[i]);
    }
    cudaFree(hist);
    return KCRN_OK;
}

}  // namespace kcrn
