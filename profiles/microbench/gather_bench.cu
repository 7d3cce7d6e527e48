// Micro-benchmark behind DESIGN.md's "gather floor": how many random 8-byte gathers per clock one SM sustains
// from (a) global memory through L1TEX (vector L2-resident), (b) its own shared memory, (c) the distributed
// shared memory of a thread-block cluster, and what it costs to fill shared memory with the vector.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu && ./gather_bench
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int kThreads = 512;

// idx: [grid][per_cta] random indices < len
__global__ void __launch_bounds__(kThreads) gather_global(const int *__restrict__ idx, const double *__restrict__ v,
                                                          int per_cta, double *out) {
    const int *my = idx + (size_t)blockIdx.x * per_cta;
    double acc = 0.0;
    for (int i = threadIdx.x; i < per_cta; i += kThreads * 4) {
        int a = my[i], b = my[i + kThreads], c = my[i + 2 * kThreads], d = my[i + 3 * kThreads];
        acc += __ldg(v + a) + __ldg(v + b) + __ldg(v + c) + __ldg(v + d);
    }
    if (acc == 1.2345) out[0] = acc;
}

__global__ void __launch_bounds__(kThreads) gather_smem(const int *__restrict__ idx, const double *__restrict__ v, int len,
                                                        int per_cta, double *out, int do_gather) {
    extern __shared__ double sv[];
    for (int i = threadIdx.x * 2; i < len; i += kThreads * 2) {   // 16-byte fill
        const double2 t = *reinterpret_cast<const double2 *>(v + i);
        sv[i] = t.x; sv[i + 1] = t.y;
    }
    __syncthreads();
    if (!do_gather) { if (sv[threadIdx.x] == 1.2345) out[0] = 1; return; }
    const int *my = idx + (size_t)blockIdx.x * per_cta;
    double acc = 0.0;
    for (int i = threadIdx.x; i < per_cta; i += kThreads * 4) {
        int a = my[i], b = my[i + kThreads], c = my[i + 2 * kThreads], d = my[i + 3 * kThreads];
        acc += sv[a] + sv[b] + sv[c] + sv[d];
    }
    if (acc == 1.2345) out[0] = acc;
}

// cluster of CS CTAs; the vector is split into CS slices of `slice` doubles, slice r lives in CTA rank r
template <int CS>
__global__ void __launch_bounds__(kThreads) gather_dsmem(const int *__restrict__ idx, const double *__restrict__ v, int slice,
                                                         int per_cta, double *out, int do_gather) {
    extern __shared__ double sv[];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    const double *mine = v + (size_t)rank * slice;
    for (int i = threadIdx.x * 2; i < slice; i += kThreads * 2) {
        const double2 t = *reinterpret_cast<const double2 *>(mine + i);
        sv[i] = t.x; sv[i + 1] = t.y;
    }
    cluster.sync();
    if (do_gather) {
        const int *my = idx + (size_t)blockIdx.x * per_cta;
        double acc = 0.0;
        for (int i = threadIdx.x; i < per_cta; i += kThreads * 4) {
            int q[4] = {my[i], my[i + kThreads], my[i + 2 * kThreads], my[i + 3 * kThreads]};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int r = q[k] / slice, o = q[k] - r * slice;
                const double *remote = cluster.map_shared_rank(sv, r);
                acc += remote[o];
            }
        }
        if (acc == 1.2345) out[0] = acc;
    }
    cluster.sync();
}

template <class F>
float time_it(F f, int reps) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); f();
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    CK(cudaGetLastError());
    return ms * 1000.f / reps;
}

int main() {
    const int grid = 148 - 148 % 8;  // 144: multiple of every cluster size tried
    const int per_cta = kThreads * 4 * 8;  // 16384 gathers per CTA (~ rcv1: 10.1k per SM per sweep)
    double *out; CK(cudaMalloc(&out, 8));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const double mhz = prop.clockRate / 1000.0;
    printf("device %s, %d SMs, clock %.0f MHz; %d CTAs x %d gathers\n", prop.name, prop.multiProcessorCount, mhz, grid, per_cta);
    for (int len : {20480, 47104, 262144, 1048576 + 262144}) {
        std::vector<int> h((size_t)grid * per_cta);
        srand(1);
        for (auto &x : h) x = (int)(((unsigned)rand() * 2654435761u) % (unsigned)len);
        int *idx; double *v;
        CK(cudaMalloc(&idx, h.size() * 4)); CK(cudaMalloc(&v, (size_t)len * 8));
        CK(cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(v, 0, (size_t)len * 8));
        float us = time_it([&] { gather_global<<<grid, kThreads>>>(idx, v, per_cta, out); }, 50);
        printf("len %8d (%6.0f KB)  global/L1TEX : %7.2f us  -> %.3f cyc/gather/SM\n", len, len * 8 / 1024.0, us,
               us * mhz / per_cta);
        if (len * 8 <= 200 * 1024) {
            CK(cudaFuncSetAttribute(gather_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, len * 8));
            float f = time_it([&] { gather_smem<<<grid, kThreads, len * 8>>>(idx, v, len, per_cta, out, 0); }, 50);
            float g = time_it([&] { gather_smem<<<grid, kThreads, len * 8>>>(idx, v, len, per_cta, out, 1); }, 50);
            printf("                       smem local   : fill %6.2f us, fill+gather %6.2f us -> %.3f cyc/gather/SM\n", f, g,
                   (g - f) * mhz / per_cta);
        }
        auto run_cluster = [&](auto kern, int cs) {
            const int slice = (len + cs - 1) / cs;
            if ((size_t)slice * 8 > 200 * 1024 || slice % 2) return;
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, slice * 8));
            if (cs > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(grid - grid % cs); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = slice * 8;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            float f = time_it([&] { CK(cudaLaunchKernelEx(&cfg, kern, (const int *)idx, (const double *)v, slice, per_cta, out, 0)); }, 50);
            float g = time_it([&] { CK(cudaLaunchKernelEx(&cfg, kern, (const int *)idx, (const double *)v, slice, per_cta, out, 1)); }, 50);
            printf("                       dsmem cluster %2d: fill %6.2f us, fill+gather %6.2f us -> %.3f cyc/gather/SM\n", cs, f, g,
                   (g - f) * mhz / per_cta);
        };
        run_cluster(gather_dsmem<2>, 2);
        run_cluster(gather_dsmem<4>, 4);
        run_cluster(gather_dsmem<8>, 8);
        run_cluster(gather_dsmem<16>, 16);
        CK(cudaFree(idx)); CK(cudaFree(v));
    }
    return 0;
}
