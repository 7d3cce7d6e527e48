// K7: the m x m cubic sub-problem, solved on the device by one CTA.
//
//   argmin_s  <g, s> + 1/2 s' T s + M/3 |s|^3 ,   T = tridiag(betas, alphas, betas),  g = |grad| e1
//
// Follows cubic_solver_root (optimizer/cubic.py:40-75): Newton's method on the secular equation
//   phi(lam) = lam^2 - M^2 |s_lam|^2 ,  s_lam = -(T + lam I)^-1 g ,
//   phi'(lam) = 2 lam + 2 M^2 s_lam' (T + lam I)^-1 s_lam ,
// iterated exactly as scipy's scalar `newton` does for root_scalar(method='newton', x0=r0,
// xtol=eps, maxiter=it_max) (scipy/optimize/_zeros_py.py:319-358): stop when phi == 0 or when
// |p - p0| <= eps, report the number of iterations the same way.  The reference solves with LAPACK
// dposv on the dense T + lam I; for a tridiagonal matrix Cholesky touches only the two diagonals, so
// the same factorisation is done here in O(m) per solve, and a non-positive pivot is reported like
// dposv's info > 0 (status = KCRN_E_NOTPD).
//
// m <= 512 and a handful of Newton iterations: the work is a chain of dependent divisions and square
// roots, so one thread walks it while the rest of the warp only helps to load and store.  The
// divisions by the Cholesky diagonal use its precomputed IEEE reciprocal plus one fused correction of
// the quotient (kcrn_device.cuh: quotient), which cuts the dependent chain of every solve step from a
// full software division to three fused operations.
#include "kcrn_device.cuh"
#include "kcrn_internal.cuh"

namespace kcrn {

namespace {

// The arrays live in shared memory and are handed to the helpers as __restrict__ pointers: the compiler may then fetch
// a[j + 1], b[j + 1], the factor's entries ... ahead of the dependent chain instead of serialising every load behind
// the previous step's store.  The arithmetic stays the Cholesky sequence of LAPACK's dpotf2 / dpotrs on the two
// diagonals (what the reference's dposv performs on this matrix): an L D L' variant with one division per row was
// measured 3x faster but rounds differently, and the chaotic tails of two trajectory cases then left the 10x envelope
// (profiles/r2i_pytest.log) -- parity first.

// Cholesky of T + lam I fused with the forward substitution of (T + lam I) y = gnorm e1 (two independent chains in one
// loop).  False when a pivot is not positive.
__device__ __forceinline__ bool factor_forward(const double *__restrict__ a, const double *__restrict__ b,
                                               double *__restrict__ l, double *__restrict__ li, double *__restrict__ e,
                                               double *__restrict__ y, int m, double lam, double gnorm) {
    double prev_e = 0.0, prev_y = 0.0;
    for (int j = 0; j < m; ++j) {
        double piv = a[j] + lam;
        piv = __dsub_rn(piv, __dmul_rn(prev_e, prev_e));
        if (!(piv > 0.0)) return false;
        const double lj = sqrt(piv);
        const double rj = 1.0 / lj;
        l[j] = lj;
        li[j] = rj;
        double v = j == 0 ? gnorm : 0.0;
        if (j > 0) v = __dsub_rn(v, __dmul_rn(prev_e, prev_y));     // e[j - 1] * y[j - 1]
        prev_y = quotient(v, lj, rj);
        y[j] = prev_y;
        if (j + 1 < m) {
            prev_e = quotient(b[j], lj, rj);
            e[j] = prev_e;
        }
    }
    return true;
}

// forward substitution with a general right-hand side
__device__ __forceinline__ void forward(const double *__restrict__ l, const double *__restrict__ li,
                                        const double *__restrict__ e, const double *__restrict__ rhs,
                                        double *__restrict__ out, int m) {
    double prev = 0.0;
    for (int j = 0; j < m; ++j) {
        double v = rhs[j];
        if (j > 0) v = __dsub_rn(v, __dmul_rn(e[j - 1], prev));
        prev = quotient(v, l[j], li[j]);
        out[j] = prev;
    }
}

// backward substitution in place; negate: the result is stored with its sign flipped (s = -(T + lam I)^-1 g)
__device__ __forceinline__ void backward(const double *__restrict__ l, const double *__restrict__ li,
                                         const double *__restrict__ e, double *__restrict__ x, int m, bool negate) {
    double prev = 0.0;
    for (int j = m - 1; j >= 0; --j) {
        double v = x[j];
        if (j + 1 < m) v = __dsub_rn(v, __dmul_rn(e[j], prev));
        prev = quotient(v, l[j], li[j]);
        x[j] = negate ? -prev : prev;
    }
}

__device__ __forceinline__ double sumsq(const double *__restrict__ x, int m) {
    double acc = 0.0;
    for (int j = 0; j < m; ++j) acc = __dadd_rn(acc, __dmul_rn(x[j], x[j]));
    return acc;
}

// r/2 |s|^2 - M/3 |s|^3 - <g, s>/2 with g = gnorm e1 (cubic.py:74): every operation rounded on its own, as numpy does
__device__ __forceinline__ double model_decrease(double r, double M, double ns, double gnorm, double s0) {
    const double n2 = __dmul_rn(ns, ns), n3 = __dmul_rn(n2, ns);
    const double t0 = __dmul_rn(r / 2.0, n2), t1 = __dmul_rn(M / 3.0, n3), t2 = __dmul_rn(gnorm, s0) / 2.0;
    return __dsub_rn(__dsub_rn(t0, t1), t2);
}

// ---- m <= kSmallM: the same arithmetic with the factor and the two solution vectors in REGISTERS.  The general kernel
// below keeps them in shared memory behind run-time loop bounds, so every step of its dependent chains goes through
// a store and a load of the same address; here the loops are unrolled over kSmallM with `j < m` guards, the factor
// never leaves the register file and the chain is arithmetic only (sqrt -> reciprocal -> quotient per row).  Same
// operations in the same order: bit-identical results (test_cubic_small_path_equals_general_path).
constexpr int kSmallM = 16;

struct SmallTri {
    double l[kSmallM], li[kSmallM], e[kSmallM];
};

__device__ __forceinline__ bool factor_forward_small(const double *__restrict__ a, const double *__restrict__ b, SmallTri &F,
                                                     double (&y)[kSmallM], int m, double lam, double gnorm) {
    double prev_e = 0.0, prev_y = 0.0;
    bool ok = true;
#pragma unroll
    for (int j = 0; j < kSmallM; ++j) {
        if (j < m) {
            double piv = a[j] + lam;
            piv = __dsub_rn(piv, __dmul_rn(prev_e, prev_e));
            if (!(piv > 0.0)) ok = false;
            const double lj = sqrt(piv);
            const double rj = 1.0 / lj;
            F.l[j] = lj;
            F.li[j] = rj;
            double v = j == 0 ? gnorm : 0.0;
            if (j > 0) v = __dsub_rn(v, __dmul_rn(prev_e, prev_y));
            prev_y = quotient(v, lj, rj);
            y[j] = prev_y;
            if (j + 1 < m) {
                prev_e = quotient(b[j], lj, rj);
                F.e[j] = prev_e;
            }
        }
    }
    return ok;
}

__device__ __forceinline__ void forward_small(const SmallTri &F, const double (&rhs)[kSmallM], double (&out)[kSmallM], int m) {
    double prev = 0.0;
#pragma unroll
    for (int j = 0; j < kSmallM; ++j) {
        if (j < m) {
            double v = rhs[j];
            if (j > 0) v = __dsub_rn(v, __dmul_rn(F.e[j - 1], prev));
            prev = quotient(v, F.l[j], F.li[j]);
            out[j] = prev;
        }
    }
}

__device__ __forceinline__ void backward_small(const SmallTri &F, double (&x)[kSmallM], int m, bool negate) {
    double prev = 0.0;
#pragma unroll
    for (int j = kSmallM - 1; j >= 0; --j) {
        if (j < m) {
            double v = x[j];
            if (j + 1 < m) v = __dsub_rn(v, __dmul_rn(F.e[j], prev));
            prev = quotient(v, F.l[j], F.li[j]);
            x[j] = negate ? -prev : prev;
        }
    }
}

__device__ __forceinline__ double dot_small(const double (&x)[kSmallM], const double (&y)[kSmallM], int m) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < kSmallM; ++j)
        if (j < m) acc = __dadd_rn(acc, __dmul_rn(x[j], y[j]));
    return acc;
}

__global__ void __launch_bounds__(32) cubic_small_kernel(const double *__restrict__ alphas,
                                                         const double *__restrict__ betas, int m_arg,
                                                         const int32_t *__restrict__ info,
                                                         const double *__restrict__ gnorm_dev, double gnorm_arg,
                                                         double M, double r0, double eps, int it_max,
                                                         double *__restrict__ s_out, double *__restrict__ scal_out,
                                                         const double *__restrict__ params) {
    __shared__ double sa[kSmallM], sb[kSmallM], sres[kSmallM];
    // all inputs are requested at once (one round trip), by different lanes
    double pv = 0.0;
    if (params && threadIdx.x < 4) pv = params[threadIdx.x];
    // (the loads are guarded by m_arg, which is a kernel argument, not by the device-side m: one round trip, not two)
    double la = 0.0, lb = 0.0;
    if ((int)threadIdx.x < m_arg) la = alphas[threadIdx.x];
    if ((int)threadIdx.x + 1 < m_arg) lb = betas[threadIdx.x];
    const int m = info ? info[0] : m_arg;
    const double gnorm = gnorm_dev ? gnorm_dev[0] : gnorm_arg;
    if (threadIdx.x < kSmallM) {
        const int j = threadIdx.x;
        sa[j] = j < m ? la : 0.0;
        sb[j] = j + 1 < m ? lb : 0.0;
    }
    if (params) {                                  // whole-step graph
        M = __shfl_sync(0xffffffffu, pv, 0); r0 = __shfl_sync(0xffffffffu, pv, 1);
        eps = __shfl_sync(0xffffffffu, pv, 2); it_max = (int)__shfl_sync(0xffffffffu, pv, 3);
    }
    __syncwarp();
    if (threadIdx.x == 0) {
        SmallTri F;
        double ss[kSmallM], sq[kSmallM];
        double r = r0, dec = 0.0;
        int its = 0, status = 0;
        double p0 = r0, p = r0;
        bool finished = false;
        for (int itr = 0; itr < it_max && !finished; ++itr) {
            if (!factor_forward_small(sa, sb, F, ss, m, p0, gnorm)) { status = KCRN_E_NOTPD; r = p0; its = itr; finished = true; break; }
            backward_small(F, ss, m, true);
            const double ns = sqrt(dot_small(ss, ss, m));
            const double fval = __dsub_rn(__dmul_rn(p0, p0), __dmul_rn(__dmul_rn(M, M), __dmul_rn(ns, ns)));
            if (fval == 0.0) { r = p0; its = itr; finished = true; break; }
            forward_small(F, ss, sq, m);
            backward_small(F, sq, m, false);
            const double sdot = dot_small(ss, sq, m);
            const double fder = __dsub_rn(__dmul_rn(2.0, p0), __dmul_rn(__dmul_rn(M, M), __dmul_rn(-2.0, sdot)));
            if (fder == 0.0) { r = p0; its = itr + 1; finished = true; break; }
            p = __dsub_rn(p0, fval / fder);
            if (fabs(p - p0) <= eps) { r = p; its = itr + 1; finished = true; break; }
            p0 = p;
        }
        if (!finished) { r = p; its = it_max; }
        if (status == 0 && !factor_forward_small(sa, sb, F, ss, m, r, gnorm)) status = KCRN_E_NOTPD;
        if (status == 0) {
            backward_small(F, ss, m, true);
            const double ns = sqrt(dot_small(ss, ss, m));
            dec = model_decrease(r, M, ns, gnorm, ss[0]);
        }
#pragma unroll
        for (int j = 0; j < kSmallM; ++j) sres[j] = (status == 0 && j < m) ? ss[j] : 0.0;
        scal_out[0] = r;
        scal_out[1] = dec;
        scal_out[2] = (double)its;
        scal_out[3] = (double)status;
    }
    __syncwarp();
    for (int j = threadIdx.x; j < m_arg; j += 32) s_out[j] = sres[j];
}

__global__ void __launch_bounds__(32) cubic_kernel(const double *__restrict__ alphas,
                                                   const double *__restrict__ betas, int m_arg,
                                                   const int32_t *__restrict__ info,
                                                   const double *__restrict__ gnorm_dev, double gnorm_arg,
                                                   double M, double r0, double eps, int it_max,
                                                   double *__restrict__ s_out, double *__restrict__ scal_out,
                                                   const double *__restrict__ params) {
    if (params) { M = params[0]; r0 = params[1]; eps = params[2]; it_max = (int)params[3]; }   // whole-step graph
    __shared__ double sa[kMaxM], sb[kMaxM], sl[kMaxM], se[kMaxM], ss[kMaxM], sq[kMaxM], sli[kMaxM];
    const int m = info ? info[0] : m_arg;
    const double gnorm = gnorm_dev ? gnorm_dev[0] : gnorm_arg;
    for (int j = threadIdx.x; j < m; j += 32) {
        sa[j] = alphas[j];
        sb[j] = j + 1 < m ? betas[j] : 0.0;
    }
    __syncwarp();
    double r = r0, dec = 0.0;
    int its = 0, status = 0;
    if (threadIdx.x == 0) {
        double p0 = r0, p = r0;
        bool finished = false;
        for (int itr = 0; itr < it_max && !finished; ++itr) {
            // s = -(T + p0 I)^-1 g with g = gnorm e1
            if (!factor_forward(sa, sb, sl, sli, se, ss, m, p0, gnorm)) { status = KCRN_E_NOTPD; r = p0; its = itr; finished = true; break; }
            backward(sl, sli, se, ss, m, true);
            const double ns = sqrt(sumsq(ss, m));
            const double fval = __dsub_rn(__dmul_rn(p0, p0), __dmul_rn(__dmul_rn(M, M), __dmul_rn(ns, ns)));
            if (fval == 0.0) { r = p0; its = itr; finished = true; break; }
            forward(sl, sli, se, ss, sq, m);
            backward(sl, sli, se, sq, m, false);
            double sdot = 0.0;
            for (int j = 0; j < m; ++j) sdot = __dadd_rn(sdot, __dmul_rn(ss[j], sq[j]));
            // 2 lam - M^2 * (-2 s.q)
            const double fder = __dsub_rn(__dmul_rn(2.0, p0), __dmul_rn(__dmul_rn(M, M), __dmul_rn(-2.0, sdot)));
            if (fder == 0.0) { r = p0; its = itr + 1; finished = true; break; }
            p = __dsub_rn(p0, fval / fder);
            if (fabs(p - p0) <= eps) { r = p; its = itr + 1; finished = true; break; }
            p0 = p;
        }
        if (!finished) { r = p; its = it_max; }
        if (status == 0) {
            if (!factor_forward(sa, sb, sl, sli, se, ss, m, r, gnorm)) status = KCRN_E_NOTPD;
        }
        if (status == 0) {
            backward(sl, sli, se, ss, m, true);
            const double ns = sqrt(sumsq(ss, m));
            // r/2 |s|^2 - M/3 |s|^3 - <g, s>/2
            dec = model_decrease(r, M, ns, gnorm, ss[0]);
        } else {
            for (int j = 0; j < m; ++j) ss[j] = 0.0;
        }
    }
    __syncwarp();
    for (int j = threadIdx.x; j < m_arg; j += 32) s_out[j] = j < m ? ss[j] : 0.0;
    if (threadIdx.x == 0) {
        scal_out[0] = r;
        scal_out[1] = dec;
        scal_out[2] = (double)its;
        scal_out[3] = (double)status;
    }
}

// ----------------------------------------------------------------------------- dense blocks (SSCN, config 4)
// The same Newton iteration for a DENSE symmetric m x m block (cubic.py:40-75 as SSCN calls it, :374-389: general g,
// epsilon = machine eps), m <= 64: the whole loop -- Cholesky of H + lam I (left-looking, column by column), the two
// triangular solves per evaluation, the norms and the stopping rule -- runs in ONE kernel of 64 threads on data in
// shared memory.  Round 1 drove it from Python with a cuSOLVER factorisation and three host round trips per
// evaluation: 2-3 ms per SSCN step, slower than the reference's CPU at m = 10 and 50.
constexpr int kDenseMax = 64;

struct DenseWork {
    const double *H;                 // the block in global memory (row-major, read-only: L1/L2 serve the re-reads)
    double (*L)[kDenseMax + 1];
    double *col;
};

// L L' = A + lam I.  Thread i owns row i.  False (in every thread) when a pivot is not positive.
__device__ bool dense_factor(const DenseWork &w, int m, double lam) {
    const int i = threadIdx.x;
    __shared__ int s_bad;
    if (i == 0) s_bad = 0;
    for (int j = 0; j < m; ++j) {
        if (i >= j && i < m) {
            double dot = 0.0;
            for (int k = 0; k < j; ++k) dot = fma(w.L[i][k], w.L[j][k], dot);
            w.col[i] = (__ldg(w.H + (size_t)i * m + j) + (i == j ? lam : 0.0)) - dot;
        }
        __syncthreads();
        const double pj = w.col[j];
        if (!(pj > 0.0)) { if (i == 0) s_bad = 1; break; }            // uniform: every thread reads the same pj
        const double ljj = sqrt(pj);
        if (i == j) w.L[j][j] = ljj;
        else if (i > j && i < m) w.L[i][j] = w.col[i] / ljj;
        __syncthreads();
    }
    __syncthreads();
    return s_bad == 0;
}

// x = (L L')^-1 rhs, in place in x (x holds rhs on entry).  Column-oriented substitutions, thread i owns x[i].
__device__ void dense_solve(const DenseWork &w, int m, double *x) {
    const int i = threadIdx.x;
    for (int j = 0; j < m; ++j) {                    // L y = rhs
        if (i == j) x[j] = x[j] / w.L[j][j];
        __syncthreads();
        if (i > j && i < m) x[i] = fma(-w.L[i][j], x[j], x[i]);
        __syncthreads();
    }
    for (int j = m - 1; j >= 0; --j) {               // L' x = y
        if (i == j) x[j] = x[j] / w.L[j][j];
        __syncthreads();
        if (i < j) x[i] = fma(-w.L[j][i], x[j], x[i]);
        __syncthreads();
    }
}

// fixed-order sum over the m entries (thread 0 walks them: m <= 64)
__device__ double dense_dot(const double *a, const double *b, int m) {
    double acc = 0.0;
    for (int k = 0; k < m; ++k) acc = __dadd_rn(acc, __dmul_rn(a[k], b[k]));
    return acc;
}

__global__ void __launch_bounds__(kDenseMax) cubic_dense_kernel(const double *__restrict__ H, const double *__restrict__ g,
                                                                int m, double M, double r0, double eps, int it_max,
                                                                double *__restrict__ out) {
    __shared__ double sL[kDenseMax][kDenseMax + 1];
    __shared__ double scol[kDenseMax], sg[kDenseMax], ss[kDenseMax], sq[kDenseMax];
    __shared__ double s_p0, s_p;
    __shared__ int s_state;                          // 0: iterate, 1: converged, 2: not positive definite
    const int i = threadIdx.x;
    DenseWork w{H, sL, scol};
    if (i < m) sg[i] = g[i];
    if (i == 0) { s_p0 = r0; s_p = r0; s_state = 0; }
    __syncthreads();
    int its = it_max;
    for (int itr = 0; itr < it_max; ++itr) {
        const double p0 = s_p0;
        if (!dense_factor(w, m, p0)) { if (i == 0) { s_state = 2; s_p = p0; } its = itr; __syncthreads(); break; }
        if (i < m) ss[i] = -sg[i];
        __syncthreads();
        dense_solve(w, m, ss);                       // s = -(H + p0 I)^-1 g
        if (i < m) sq[i] = ss[i];
        __syncthreads();
        dense_solve(w, m, sq);                       // (H + p0 I)^-1 s
        if (i == 0) {
            const double ns = sqrt(dense_dot(ss, ss, m));
            const double fval = p0 * p0 - M * M * ns * ns;
            if (fval == 0.0) { s_p = p0; s_state = 1; s_p0 = (double)itr; }
            else {
                const double fder = 2.0 * p0 + 2.0 * M * M * dense_dot(ss, sq, m);
                if (fder == 0.0) { s_p = p0; s_state = 1; s_p0 = (double)(itr + 1); }
                else {
                    const double p = p0 - fval / fder;
                    s_p = p;
                    if (fabs(p - p0) <= eps) { s_state = 1; s_p0 = (double)(itr + 1); }
                    else s_p0 = p;
                }
            }
        }
        __syncthreads();
        if (s_state == 1) { its = (int)s_p0; break; }
    }
    __syncthreads();
    const double r = s_p;
    double dec = 0.0;
    int status = s_state == 2 ? KCRN_E_NOTPD : 0;
    if (status == 0 && !dense_factor(w, m, r)) status = KCRN_E_NOTPD;
    if (status == 0) {
        if (i < m) ss[i] = -sg[i];
        __syncthreads();
        dense_solve(w, m, ss);
        if (i == 0) {
            const double ns = sqrt(dense_dot(ss, ss, m));
            dec = r / 2.0 * (ns * ns) - M / 3.0 * (ns * ns * ns) - dense_dot(sg, ss, m) / 2.0;
        }
    } else if (i < m) {
        ss[i] = 0.0;
    }
    __syncthreads();
    if (i < m) out[i] = ss[i];
    if (i == 0) { out[m] = r; out[m + 1] = dec; out[m + 2] = (double)its; out[m + 3] = (double)status; }
}

}  // namespace

int launch_cubic_dense(const double *H, const double *g, int m, double M, double r0, double eps, int it_max, double *out,
                       cudaStream_t st) {
    if (m < 1 || m > kDenseMax) {
        set_error("dense cubic sub-problem: m = %d outside [1, %d]", m, kDenseMax);
        return KCRN_E_ARG;
    }
    cubic_dense_kernel<<<1, kDenseMax, 0, st>>>(H, g, m, M, r0, eps, it_max, out);
    KCRN_LAUNCH_CHECK();
    return KCRN_OK;
}

int launch_cubic(const double *alphas, const double *betas, int m, const int32_t *info, const double *gnorm_dev,
                 double gnorm, double M, double r0, double eps, int it_max, double *s_out, double *scal_out,
                 cudaStream_t st, const double *params_dev) {
    if (m < 1 || m > kMaxM) {
        set_error("cubic sub-problem: m = %d outside [1, %d]", m, kMaxM);
        return KCRN_E_ARG;
    }
    if (m <= kSmallM)
        cubic_small_kernel<<<1, 32, 0, st>>>(alphas, betas, m, info, gnorm_dev, gnorm, M, r0, eps, it_max, s_out, scal_out,
                                             params_dev);
    else
        cubic_kernel<<<1, 32, 0, st>>>(alphas, betas, m, info, gnorm_dev, gnorm, M, r0, eps, it_max, s_out, scal_out,
                                       params_dev);
    KCRN_LAUNCH_CHECK();
    return KCRN_OK;
}

}  // namespace kcrn
