// Device-side helpers shared by the kernels of libkcrn.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace kcrn {

// Fixed butterfly: every lane ends with the same total, the order of additions never changes.
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming loads of the matrix arrays: read-only path, do not keep the line in L1
__device__ __forceinline__ int ld_stream(const int32_t *p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_stream(const double *p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// 4-byte asynchronous copy global -> shared (LDGSTS); completion with cp_async_wait_all()
__device__ __forceinline__ void cp_async_4(void *smem_dst, const void *gmem_src) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(gmem_src) : "memory");
}
// 8-byte variant
__device__ __forceinline__ void cp_async_8(void *smem_dst, const void *gmem_src) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(gmem_src) : "memory");
}
// 16-byte variant, cached in L2 only (.cg)
__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gmem_src) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// groups: commit what was issued so far; wait until at most the most recent group is still in flight
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_group_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// ---- mbarrier + bulk copy (cp.async.bulk, SASS UBLKCP): one thread asks for a whole contiguous block of global memory;
// it arrives in shared memory without passing through registers or the L1TEX request queue
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy accesses to shared memory issued so far are ordered before later async-proxy (bulk copy) writes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy (16-byte aligned, size a multiple of 16); completion is counted on the mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "PWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra PDONE_%=;\n"
        "bra PWAIT_%=;\n"
        "PDONE_%=:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// x / b with b fixed for the whole kernel: rinv = 1/b (IEEE) is formed once per thread, then one
// Newton correction on the quotient (Markstein).  Bit-identical for every reader of the same x, so
// the stored basis vector and the values gathered on the fly agree exactly.
__device__ __forceinline__ double quotient(double x, double b, double rinv) {
    const double q = x * rinv;
    const double r = fma(-b, q, x);
    return fma(r, rinv, q);
}

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-serialisation attribute may
// become resident while its predecessor in the stream is still running.  pdl_wait() blocks until that
// predecessor has completed and its writes are visible (no-op for ordinary launches); pdl_trigger() lets the
// successor start.  Every kernel of this library waits before it touches anything a previous kernel wrote and
// only triggers afterwards, so completion stays transitive along the stream.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// scipy.special.expit as the reference calls it (optimizer/loss.py:225,296): 1 / (1 + exp(-x))
__device__ __forceinline__ double sigmoid(double x) { return 1.0 / (1.0 + exp(-x)); }

// log(sigmoid(x)) in the four ranges of the reference's jitted helper (optimizer/loss.py:161-176)
__device__ __forceinline__ double log_sigmoid(double x) {
    if (x < -33.0) return x;
    if (x < -18.0) return x - exp(x);
    if (x < 37.0) return -log1p(exp(-x));
    return -exp(-x);
}

}  // namespace kcrn
