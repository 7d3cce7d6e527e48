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
#include <stdlib.h>
#include <string.h>

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

// `map` (first pass only): sort by map[key] instead of key -- the position of the column in the sweep order
__global__ void __launch_bounds__(kSortThreads) radix_hist(const int32_t *__restrict__ keys, int64_t n,
                                                           int shift, int32_t *__restrict__ hist, int nblocks,
                                                           const int32_t *__restrict__ map) {
    __shared__ int sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
    for (int i = threadIdx.x; i < kSortTile; i += kSortThreads) {
        const int64_t p = base + i;
        if (p < n) {
            const int key = map ? map[keys[p]] : keys[p];
            atomicAdd(&sh[(key >> shift) & 255], 1);  // integer counts: order-free
        }
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = sh[threadIdx.x];
}

// Stable scatter.  Order inside the CTA is (warp, round, lane) = ascending input position.
__global__ void __launch_bounds__(kSortThreads) radix_scatter(const int32_t *__restrict__ keys_in,
                                                              const int32_t *__restrict__ pay_in, int64_t n,
                                                              int shift, const int32_t *__restrict__ offs,
                                                              int nblocks, int32_t *__restrict__ keys_out,
                                                              int32_t *__restrict__ pay_out,
                                                              const int32_t *__restrict__ map) {
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
            const int dg = ((map ? map[keys_in[p]] : keys_in[p]) >> shift) & 255;
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
            const int key = map ? map[keys_in[p]] : keys_in[p];
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

// column lengths (integer atomics: order-free)
__global__ void column_counts(const int32_t *__restrict__ idx, int64_t nnz, int32_t *__restrict__ counts) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(counts + idx[k], 1);
}

// canonical (scipy) column order out of the sweep order: one warp copies one column
__global__ void canonical_columns(const int32_t *__restrict__ sptr, const int32_t *__restrict__ sidx,
                                  const double *__restrict__ sval, const int32_t *__restrict__ pos,
                                  const int32_t *__restrict__ cptr, int64_t n_cols, int32_t *__restrict__ oidx,
                                  double *__restrict__ oval) {
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= n_cols) return;
    const int src = sptr[pos[c]], dst = cptr[c], len = cptr[c + 1] - dst;
    for (int i = lane; i < len; i += 32) { oidx[dst + i] = sidx[src + i]; oval[dst + i] = sval[src + i]; }
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

int build_csc(const SegMat &csr, int64_t n_cols, SegMat *csc, cudaStream_t s, bool natural_order) {
    const int64_t nnz = csr.nnz, n_rows = csr.nseg;
    csc->nseg = n_cols;
    csc->nnz = nnz;
    KCRN_CUDA(cudaMalloc(&csc->ptr, sizeof(int32_t) * (n_cols + 1)));
    KCRN_CUDA(cudaMalloc(&csc->idx, sizeof(int32_t) * (nnz > 0 ? nnz : 1)));
    KCRN_CUDA(cudaMalloc(&csc->val, sizeof(double) * (nnz > 0 ? nnz : 1)));
    if (nnz == 0) {
        KCRN_CUDA(cudaMemsetAsync(csc->ptr, 0, sizeof(int32_t) * (n_cols + 1), s));
        csc->cptr_host.assign((size_t)n_cols + 1, 0);
        return KCRN_OK;
    }
    // Sweep order of the columns: a stable partition by length class.  Long and medium columns scattered among
    // short ones would cut the runs of short columns that share a tile (the rcv1-shaped CSC needed 1 109 tiles for
    // 1.5 M non-zeros); grouped by class the tiles fill up and hold segments of similar length.  Inside a class
    // the natural order is kept, so the per-column operands of the epilogue stay mostly coalesced.
    {
        int32_t *counts = nullptr;
        KCRN_CUDA(cudaMalloc(&counts, sizeof(int32_t) * n_cols));
        KCRN_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * n_cols, s));
        column_counts<<<kNumSm * 8, 256, 0, s>>>(csr.idx, nnz, counts);
        KCRN_LAUNCH_CHECK();
        std::vector<int32_t> cnt((size_t)n_cols), pos((size_t)n_cols), order((size_t)n_cols);
        KCRN_CUDA(cudaMemcpyAsync(cnt.data(), counts, sizeof(int32_t) * n_cols, cudaMemcpyDeviceToHost, s));
        KCRN_CUDA(cudaStreamSynchronize(s));
        cudaFree(counts);
        const bool natural = natural_order;   // A/B switch: keep the scipy column order
        auto cls = [natural](int len) { return natural ? 0 : len <= 8 ? 0 : len <= 32 ? 1 : len <= 128 ? 2 : len <= 512 ? 3 : len <= kTileNnz ? 4 : 5; };
        int64_t start[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int64_t c = 0; c < n_cols; ++c) start[cls(cnt[c]) + 1] += 1;
        for (int k = 0; k < 6; ++k) start[k + 1] += start[k];
        bool identity = true;
        for (int64_t c = 0; c < n_cols; ++c) {
            const int64_t q = start[cls(cnt[c])]++;
            pos[c] = (int32_t)q;
            order[q] = (int32_t)c;
            identity = identity && q == c;
        }
        csc->cptr_host.assign((size_t)n_cols + 1, 0);
        for (int64_t c = 0; c < n_cols; ++c) csc->cptr_host[c + 1] = csc->cptr_host[c] + cnt[c];
        if (!identity) {
            KCRN_CUDA(cudaMalloc(&csc->pos, sizeof(int32_t) * n_cols));
            KCRN_CUDA(cudaMalloc(&csc->segid, sizeof(int32_t) * n_cols));
            KCRN_CUDA(cudaMemcpyAsync(csc->pos, pos.data(), sizeof(int32_t) * n_cols, cudaMemcpyHostToDevice, s));
            KCRN_CUDA(cudaMemcpyAsync(csc->segid, order.data(), sizeof(int32_t) * n_cols, cudaMemcpyHostToDevice, s));
            KCRN_CUDA(cudaStreamSynchronize(s));
        }
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
        const int32_t *map = p == 0 ? csc->pos : nullptr;
        radix_hist<<<nblocks, kSortThreads, 0, s>>>(kin, nnz, shift, hist, nblocks, map);
        KCRN_LAUNCH_CHECK();
        KCRN_TRY(exclusive_scan_i32(hist, hist, hlen, s));
        radix_scatter<<<nblocks, kSortThreads, 0, s>>>(kin, pin, nnz, shift, hist, nblocks, keys[cur], pay[cur], map);
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

// canonical CSC (column order and entry order of scipy's csr_tocsc) into host arrays; synchronous
int canonical_csc_to_host(const SegMat &csc, int32_t *indptr_host, int32_t *indices_host, double *data_host,
                          cudaStream_t s) {
    const int64_t n_cols = csc.nseg, nnz = csc.nnz;
    if (indptr_host) memcpy(indptr_host, csc.cptr_host.data(), sizeof(int32_t) * (n_cols + 1));
    if (nnz == 0 || (!indices_host && !data_host)) return KCRN_OK;
    const int32_t *idx = csc.idx;
    const double *val = csc.val;
    int32_t *cptr = nullptr, *oidx = nullptr;
    double *oval = nullptr;
    if (csc.pos) {
        KCRN_CUDA(cudaMalloc(&cptr, sizeof(int32_t) * (n_cols + 1)));
        KCRN_CUDA(cudaMalloc(&oidx, sizeof(int32_t) * nnz));
        KCRN_CUDA(cudaMalloc(&oval, sizeof(double) * nnz));
        KCRN_CUDA(cudaMemcpyAsync(cptr, csc.cptr_host.data(), sizeof(int32_t) * (n_cols + 1), cudaMemcpyHostToDevice, s));
        canonical_columns<<<(unsigned)((n_cols * 32 + 255) / 256), 256, 0, s>>>(csc.ptr, csc.idx, csc.val, csc.pos, cptr,
                                                                                n_cols, oidx, oval);
        KCRN_LAUNCH_CHECK();
        idx = oidx;
        val = oval;
    }
    if (indices_host) KCRN_CUDA(cudaMemcpyAsync(indices_host, idx, sizeof(int32_t) * nnz, cudaMemcpyDeviceToHost, s));
    if (data_host) KCRN_CUDA(cudaMemcpyAsync(data_host, val, sizeof(double) * nnz, cudaMemcpyDeviceToHost, s));
    KCRN_CUDA(cudaStreamSynchronize(s));
    cudaFree(cptr); cudaFree(oidx); cudaFree(oval);
    return KCRN_OK;
}

}  // namespace kcrn
