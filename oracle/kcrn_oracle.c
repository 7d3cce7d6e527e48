/*
 * kcrn_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C, float64/int32) of the
 * arithmetic the reference's Krylov-CRN hot path executes inside third-party libraries.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file; the shipped product path never does.
 *
 * The reference (/root/reference, pure Python) delegates the numerics to scipy 1.10.1 /
 * numpy 1.24.3 / numba 0.57.0 (requirements.txt:2-5); none of that source is under
 * /root/reference, so the published algorithms are restated here and pinned against the
 * installed scipy 1.18.1 / numpy 2.3.5 / numba 0.65.0 by tests/test_oracle_golden.py:
 *
 *   ko_csr_tocsc     scipy sparsetools csr_tocsc     <- reference cubic_newton.py:55  (A.tocsc())
 *   ko_csr_matvec    scipy sparsetools csr_matvec    <- reference optimizer/loss.py:270, :299 (A @ x)
 *   ko_csc_matvec    scipy sparsetools csc_matvec    <- reference optimizer/loss.py:227, :302 (A.T @ r)
 *   ko_cscT_gather   same sums as ko_csc_matvec, gathered per output (thread-parallel, same order)
 *   ko_expit         scipy.special.expit             <- reference optimizer/loss.py:225, :296
 *   ko_logsig        numba-jitted logsig             <- reference optimizer/loss.py:161-176
 *   ko_posv          LAPACK dposv (Cholesky solve)   <- reference optimizer/cubic.py:52,58 (solve(assume_a='pos'))
 *   ko_hvp_weights   activation*(1-activation)       <- reference optimizer/loss.py:296-297
 *
 * Parity status: the reference ships no tests or golden vectors; this oracle is pinned
 * against outputs of the reference itself run in the build container (tests/golden/make_golden.py).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef int32_t i32;
typedef int64_t i64;

int ko_version(void) { return 1; }

int ko_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Counting-sort transpose.  Column pointer = exclusive scan of the column histogram; entries are
 * dropped into their column in (row, position-in-row) order, so equal columns keep CSR order and
 * duplicates survive. */
void ko_csr_tocsc(i64 n_row, i64 n_col, const i32 *Ap, const i32 *Aj, const double *Ax,
                  i32 *Bp, i32 *Bi, double *Bx) {
    const i64 nnz = Ap[n_row];
    for (i64 c = 0; c <= n_col; ++c) Bp[c] = 0;
    for (i64 k = 0; k < nnz; ++k) Bp[Aj[k]] += 1;
    i32 run = 0;
    for (i64 c = 0; c < n_col; ++c) {
        i32 cnt = Bp[c];
        Bp[c] = run;
        run += cnt;
    }
    Bp[n_col] = (i32)nnz;
    for (i64 r = 0; r < n_row; ++r) {
        for (i32 k = Ap[r]; k < Ap[r + 1]; ++k) {
            i32 c = Aj[k];
            i32 slot = Bp[c]++;
            Bi[slot] = (i32)r;
            Bx[slot] = Ax[k];
        }
    }
    i32 prev = 0;
    for (i64 c = 0; c <= n_col; ++c) {
        i32 t = Bp[c];
        Bp[c] = prev;
        prev = t;
    }
}

/* y = A x, A in CSR; each row is one left-to-right running sum starting from 0. */
void ko_csr_matvec(i64 n_row, const i32 *Ap, const i32 *Aj, const double *Ax,
                   const double *x, double *y, int threads) {
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
    for (i64 r = 0; r < n_row; ++r) {
        double acc = 0.0;
        for (i32 k = Ap[r]; k < Ap[r + 1]; ++k) acc += Ax[k] * x[Aj[k]];
        y[r] = acc;
    }
}

/* y = A^T r with A given in CSR, computed the way scipy does it for the zero-copy CSC view of
 * A.T: sweep rows in order and scatter into y (serial by construction). */
void ko_csc_matvec(i64 n_row, i64 n_col, const i32 *Ap, const i32 *Aj, const double *Ax,
                   const double *r, double *y) {
    for (i64 c = 0; c < n_col; ++c) y[c] = 0.0;
    for (i64 i = 0; i < n_row; ++i) {
        const double ri = r[i];
        for (i32 k = Ap[i]; k < Ap[i + 1]; ++k) y[Aj[k]] += Ax[k] * ri;
    }
}

/* Same numbers as ko_csc_matvec but gathered from the CSC copy (Bp,Bi,Bx = ko_csr_tocsc(A)):
 * every y[c] adds its terms in increasing (row, position) order, i.e. the order the scatter above
 * reaches them, so the result is bit-identical while the columns can run on separate threads. */
void ko_cscT_gather(i64 n_col, const i32 *Bp, const i32 *Bi, const double *Bx,
                    const double *r, double *y, int threads) {
#pragma omp parallel for schedule(dynamic, 512) num_threads(threads) if (threads > 1)
    for (i64 c = 0; c < n_col; ++c) {
        double acc = 0.0;
        for (i32 k = Bp[c]; k < Bp[c + 1]; ++k) acc += Bx[k] * r[Bi[k]];
        y[c] = acc;
    }
}

/* logistic sigmoid, 1/(1+exp(-x)) */
void ko_expit(i64 n, const double *x, double *out) {
    for (i64 i = 0; i < n; ++i) out[i] = 1.0 / (1.0 + exp(-x[i]));
}

/* log(sigmoid(x)) in four ranges (Fabian Pedregosa's recipe, as the reference's jitted helper) */
void ko_logsig(i64 n, const double *x, double *out) {
    for (i64 i = 0; i < n; ++i) {
        const double t = x[i];
        double v;
        if (t < -33.0) v = t;
        else if (t < -18.0) v = t - exp(t);
        else if (t < 37.0) v = -log1p(exp(-t));
        else v = -exp(-t);
        out[i] = v;
    }
}

/* D = a (1 - a) with a = expit(Ax) rounded first (loss.py:296-297) */
void ko_hvp_weights(i64 n, const double *Ax, double *w) {
    for (i64 i = 0; i < n; ++i) {
        const double a = 1.0 / (1.0 + exp(-Ax[i]));
        w[i] = a * (1.0 - a);
    }
}

/* Solve S z = rhs for symmetric positive definite S (row-major m x m, lower triangle read),
 * by unblocked Cholesky + two triangular sweeps.  `work` holds m*m doubles.
 * Returns 0, or k+1 when the k-th pivot is not positive (LAPACK's info>0 -> LinAlgError). */
int ko_posv(i64 m, const double *S, const double *rhs, double *z, double *work) {
    double *L = work;
    memcpy(L, S, (size_t)(m * m) * sizeof(double));
    for (i64 j = 0; j < m; ++j) {
        double djj = L[j * m + j];
        for (i64 k = 0; k < j; ++k) djj -= L[j * m + k] * L[j * m + k];
        if (!(djj > 0.0)) return (int)(j + 1);
        djj = sqrt(djj);
        L[j * m + j] = djj;
        for (i64 i = j + 1; i < m; ++i) {
            double v = L[i * m + j];
            for (i64 k = 0; k < j; ++k) v -= L[i * m + k] * L[j * m + k];
            L[i * m + j] = v / djj;
        }
    }
    for (i64 i = 0; i < m; ++i) {
        double v = rhs[i];
        for (i64 k = 0; k < i; ++k) v -= L[i * m + k] * z[k];
        z[i] = v / L[i * m + i];
    }
    for (i64 i = m - 1; i >= 0; --i) {
        double v = z[i];
        for (i64 k = i + 1; k < m; ++k) v -= L[k * m + i] * z[k];
        z[i] = v / L[i * m + i];
    }
    return 0;
}
