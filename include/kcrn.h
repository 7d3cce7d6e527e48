/*
 * kcrn.h -- C ABI of the B200-native Krylov cubic-regularised-Newton engine (libkcrn.so).
 *
 * The reference (Raymond30/Krylov-Cubic-Regularized-Newton) is pure Python and has no FFI of its
 * own; its boundary for this path is the duck-typed loss oracle handed to every optimizer plus two
 * free functions.  Each entry point below names the reference interface it stands in for
 * (file:line in the reference tree).  INTEGRATION.md shows the ctypes stub that binds them.
 *
 * Conventions
 *   - every function returns 0 on success, a negative KCRN_E_* code otherwise;
 *     kcrn_last_error() gives the message of the last failure on the calling thread.
 *   - pointers named *_dev are device pointers (float64 / int32), *_host are host pointers.
 *   - all work is enqueued on the stream given to kcrn_create / kcrn_set_stream (a cudaStream_t
 *     passed as void*; NULL = the legacy default stream) and is asynchronous unless stated.
 *   - one handle = one row shard of the data matrix on one GPU; one process per GPU.
 *   - there is no CPU fallback: every call fails with KCRN_E_CUDA when no sm_100 device is usable.
 */
#ifndef KCRN_H
#define KCRN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KCRN_ABI_VERSION 1

#define KCRN_OK 0
#define KCRN_E_ARG (-1)     /* bad argument (shape, null pointer, unsupported size) */
#define KCRN_E_CUDA (-2)    /* CUDA runtime / launch failure */
#define KCRN_E_STATE (-3)   /* call order violated (e.g. gradient before set_x) */
#define KCRN_E_NOTPD (-4)   /* T + r I not positive definite (LAPACK dposv info > 0 in the reference) */
#define KCRN_E_NCCL (-5)    /* NCCL unavailable or failed */

typedef struct kcrn_problem *kcrn_handle;

int kcrn_abi_version(void);
const char *kcrn_last_error(void);

/* ---- problem set-up ------------------------------------------------------------------------
 * Replaces LogisticRegression.__init__ (optimizer/loss.py:186-213) and the driver's A.tocsc()
 * (cubic_newton.py:55).  Takes one CSR row shard (int32 indices, float64 values), labels already
 * mapped to {0,1} (the host mirror does loss.py:190-207), the l2 coefficient and the GLOBAL number
 * of samples (the 1/n of loss.py:220,227,302).  Copies the arrays to the device (arrays_on_device
 * = 0) or adopts device arrays by copy (= 1), builds the CSC copy on the device with a stable
 * counting sort (bit-identical to scipy's csr_tocsc) and the nnz-balanced tile lists. */
int kcrn_create(kcrn_handle *out, int64_t n_rows, int64_t n_cols, int64_t nnz,
                const int32_t *indptr, const int32_t *indices, const double *data,
                const double *labels01, double l2, int64_t n_global, int arrays_on_device,
                void *stream);
int kcrn_destroy(kcrn_handle h);
int kcrn_set_stream(kcrn_handle h, void *stream);
int kcrn_sync(kcrn_handle h);

/* Row-sharded runs: join an NCCL communicator (one rank per handle).  unique_id is the 128-byte
 * ncclUniqueId produced by kcrn_nccl_unique_id on rank 0 and broadcast by the host plumbing
 * (torch.distributed).  After this every gradient / HVP / value adds one allreduce. */
int kcrn_nccl_unique_id(void *unique_id_128_host);
int kcrn_comm_init(kcrn_handle h, const void *unique_id_128_host, int rank, int world);

/* Optional, after kcrn_comm_init on ranks that share one NVLink domain (one GPU per rank): receive buffers mapped
 * into every peer (CUDA IPC).  export writes this rank's 64-byte cudaIpcMemHandle_t; the host plumbing all-gathers
 * them (world x 64 bytes, rank order) and hands them to attach.  The fused Lanczos step then sums the partial A't
 * of all ranks inside its epilogue kernel -- each rank pushes {value, step} slots into its peers over NVLink and
 * polls its own buffer; no NCCL call -- whenever (world - 1)(d + 2) 16 bytes <= 8 MiB, i.e. where the collective is
 * latency-bound (otherwise both calls are no-ops and NCCL stays); gradient, HVP and value keep NCCL. */
int kcrn_p2p_export(kcrn_handle h, void *handle_64_host);
int kcrn_p2p_attach(kcrn_handle h, const void *all_handles_host);
/* Collective (every rank, same value): switch between the fused peer exchange (1) and NCCL (0); A/B and tests. */
int kcrn_p2p_enable(kcrn_handle h, int on);

/* Device CSC copy -> host arrays (int32[n_cols+1], int32[nnz], float64[nnz]); synchronous.
 * Lets the tests compare with A.tocsc() (cubic_newton.py:55) bit for bit. */
int kcrn_csc_view(kcrn_handle h, int32_t *indptr_host, int32_t *indices_host, double *data_host);

/* ---- loss oracle ---------------------------------------------------------------------------
 * kcrn_set_x: forward sweep at x -- caches A x, sigma(Ax), the Hessian weights a(1-a), the
 * residual a-b and the per-sample loss terms.  Replaces the value-equality memo
 * mat_vec_product (loss.py:266-277): the host mirror decides when x changed.            */
int kcrn_set_x(kcrn_handle h, const double *x_dev);
/* The iterate changes but the cached A x must stay as it is (update_mat_vec_product overrides it until reset(),
 * loss.py:279-286): only the copy of x used by the l2 terms of value / partial_gradient (loss.py:217, :247) moves. */
int kcrn_set_x_only(kcrn_handle h, const double *x_dev);
/* value at the current x -> *out_dev (1 double).  Oracle._value, loss.py:215-220 (+ logsig :161-176). */
int kcrn_value(kcrn_handle h, double *out_dev);
/* gradient at the current x -> out_dev[n_cols] (NULL: keep it in the handle for kcrn_lanczos).
 * LogisticRegression.gradient, loss.py:223-232. */
int kcrn_gradient(kcrn_handle h, double *out_dev);
/* Hessian-vector product at the current x.  LogisticRegression.hess_vec_prod, loss.py:289-302. */
int kcrn_hvp(kcrn_handle h, const double *v_dev, double *out_dev);
/* Same three through HOST buffers (pinned staging inside; synchronous) -- the end-to-end entry
 * points a numpy caller of the reference oracle would use. */
int kcrn_value_grad_host(kcrn_handle h, const double *x_host, double *value_host, double *grad_host);
int kcrn_hvp_host(kcrn_handle h, const double *x_host_or_null, const double *v_host, double *out_host);
/* Cached A x of the current x -> device pointer (n_rows doubles), for tests. */
int kcrn_cached_Ax(kcrn_handle h, const double **Ax_dev);
/* SSCN support (loss.py:279-286): overwrite the cached A x (and everything derived from it). */
int kcrn_set_Ax(kcrn_handle h, const double *Ax_dev);

/* ---- Lanczos -------------------------------------------------------------------------------
 * m-step Lanczos on v -> H(x) v started from g (NULL: the gradient kept by kcrn_gradient), the
 * whole loop enqueued as one CUDA graph without host round trips.  Lanczos(), cubic.py:77-111,
 * including its breakdown rules (:98-99, :105-109).  reorth = 0 is the reference recurrence;
 * reorth = 1 adds two classical Gram-Schmidt passes against the basis per step (not in the
 * reference; SURVEY.md section 8(c) "reorth oracle").  Results stay in the handle: */
int kcrn_lanczos(kcrn_handle h, const double *g_dev, int m, int reorth);
/* Device-resident outer loop: up to max_iters iterations of Cubic_Krylov_LS.step
 * (optimizer/cubic.py:265-309) as ONE graph launch -- the line search ("M = reg_coef * beta first", the test
 * value_new > value - model_decrease and tries < max_tries, M / beta on rejection; cubic.py:286-302), the commit of
 * (x, value, M, r0), the stopping rule |x_new - x| < tol (optimizer.py:110; tol <= 0: off) and the best-point
 * tracking of Oracle.value (loss.py:70-72) run on the device, driving CUDA graph WHILE nodes; no host round trip
 * until the run ends.  Needs kcrn_set_x; value0 = f at that x; f_opt = best value seen so far (+inf: none).
 * Out: the number of iterations done, the sub-problem status of the last trial (0, or KCRN_E_NOTPD: the run stopped
 * there), and one 16-double record per iteration: the 12 of kcrn_step_record ([11] = best value seen so far), accepted M, backtracks, seconds
 * since the start of the run (device clock) and a flag "x_opt changed during this run".  keep_xs: also keep every
 * accepted iterate (kcrn_run_view).  Afterwards the handle's x and caches are those of the last accepted iterate.
 * Row-sharded runs need the peer exchange (kcrn_p2p_attach). */
int kcrn_krylov_run(kcrn_handle h, int max_iters, int m, int reorth, double value0, double reg_coef, double beta,
                    double r0, double eps, int solver_it_max, double tol, int max_tries, int keep_xs, double f_opt,
                    int *iters_done_host, int *status_host, double *recs16_host);
/* f(x_j) for k stored iterates X[k][d] (device, row-major) -> out[k]: Trace.compute_loss_of_iterates
 * (optimizer/opt_trace.py:39-43) with the matrix read once per 8 iterates instead of once per iterate.  Leaves the
 * caches of the current x alone.  Row-sharded: every rank passes the same X, every rank gets the full values. */
int kcrn_values_batch(kcrn_handle h, const double *X_dev, int k, double *out_dev);

/* Device views after kcrn_krylov_run: accepted iterates [iters][d] (null without keep_xs), the best point [d]
 * (valid when the record's flag is set), the current iterate [d]. */
int kcrn_run_view(kcrn_handle h, const double **xs_dev, const double **x_opt_dev, const double **x_dev);
/* Copy of the handle's current iterate into out[d] (device), enqueued on the handle's stream: the `x_new` a caller of
 * Cubic_Krylov_LS.step keeps (optimizer/cubic.py:291, 303) after a one-iteration kcrn_krylov_run. */
int kcrn_current_x(kcrn_handle h, double *out_dev);

/* basis V as m contiguous d-vectors, alphas[m], bet[m] with bet[0] = |g| and bet[j+1] = beta_j,
 * info[8] int32 = {m_eff, n_valid_columns, broke, hvps, 0...}.  Any pointer may be NULL. */
int kcrn_lanczos_view(kcrn_handle h, const double **V_dev, const double **alphas_dev,
                      const double **bet_dev, const int32_t **info_dev);
/* Step-wise form for an arbitrary operator (the host calls the operator between steps):
 * begin normalises g into V[0]; feed consumes hv = A(V[cur]) and advances the recurrence;
 * cur_dev points at the vector the operator must be applied to next.  Host-driven, so each
 * feed returns whether the recurrence is finished (synchronises on one int). */
int kcrn_lanczos_begin(kcrn_handle h, const double *g_dev, int m, int reorth);
int kcrn_lanczos_current(kcrn_handle h, const double **cur_dev);
int kcrn_lanczos_feed(kcrn_handle h, const double *hv_dev, int *finished_host);

/* ---- cubic sub-problem ---------------------------------------------------------------------
 * argmin_s <g,s> + 1/2 s'Ts + M/3 |s|^3 for tridiagonal T (alphas, betas) and g = gnorm e1, by
 * Newton on the secular equation exactly as cubic_solver_root (cubic.py:40-75) drives scipy's
 * newton (x0 = r0, |p - p0| <= eps stop, it_max).  Single-CTA kernel.
 * out_dev[0..m) = s, out_dev[m..m+4) = {r, model_decrease, iterations, status}. */
int kcrn_cubic_tridiag(const double *alphas_dev, const double *betas_dev, int m, double gnorm,
                       double M, double r0, double eps, int it_max, double *out_dev, void *stream);
/* The same sub-problem for a DENSE symmetric block H[m][m] (row-major) and a general g[m], m <= 64 -- what SSCN passes
 * to cubic_solver_root (optimizer/cubic.py:374-389; reference solves with scipy.linalg.solve(assume_a='pos') = dposv).
 * One kernel: Cholesky of H + lam I, both solves, norms and the Newton loop of scipy's scalar `newton`.
 * out_dev[m + 4] = {s[0..m), r, model decrease, Newton iterations, status (0 or KCRN_E_NOTPD)}. */
int kcrn_cubic_dense(const double *H_dev, const double *g_dev, int m, double M, double r0, double eps, int it_max,
                     double *out_dev, void *stream);
/* Same, on the handle's Lanczos result (m_eff and |g| read on the device). */
int kcrn_cubic_from_lanczos(kcrn_handle h, double M, double r0, double eps, int it_max);

/* ---- trial point ---------------------------------------------------------------------------
 * out = x + V s  (cubic.py:291,301), generic form and the handle form (V, s, m_eff from the
 * handle).  */
int kcrn_axpy_basis(const double *x_dev, const double *V_dev, const double *s_dev, int m,
                    int64_t d, double *out_dev, void *stream);
int kcrn_trial_point(kcrn_handle h, const double *x_dev, double *out_dev);
/* One D2H of the step record {value, model_decrease, r, iterations, status, m_eff, |g|, hvps,
 * |x_trial - x|, basis columns written, breakdown flag, 0}; synchronous.  value is the value at
 * the current x (the value reduction is enqueued by this call when with_value != 0). */
int kcrn_step_record(kcrn_handle h, int with_value, double *rec12_host);

/* One trial of a Krylov-CRN outer iteration as ONE CUDA graph plus one device-to-host copy (cubic.py:265-303
 * without the host-side compare).  first != 0: keep the current x as the base iterate, gradient, m-step
 * Lanczos; always: sub-problem with (M, r0, eps, it_max), trial point x_base + V s, forward sweep and value
 * at the trial point (which becomes the handle's current x).  first == 0 repeats only the second half with a
 * new M (the backtracking loop of cubic.py:294-303).  x_out_dev (nullable) receives the trial point;
 * rec12_host as in kcrn_step_record.  Synchronous. */
int kcrn_krylov_trial(kcrn_handle h, int first, int m, int reorth, double M, double r0, double eps, int it_max,
                      double *x_out_dev, double *rec12_host);

/* ---- config-4 rows (full-Hessian CRN, SSCN) -------------------------------------------------
 * dense A' D A / n + l2 I, row-major [n_cols x n_cols]  (loss.py:249-255), FP64 tensor cores. */
int kcrn_hessian_dense(kcrn_handle h, double *out_dev);
/* block gradient / Hessian over the column set I (loss.py:234-247, :257-264) and the
 * incremental A x update  Ax + A[:, I] delta  (loss.py:279-281). */
int kcrn_partial_gradient(kcrn_handle h, const int64_t *I_dev, int m, double *out_dev);
int kcrn_partial_hessian(kcrn_handle h, const int64_t *I_dev, int m, double *out_dev);
int kcrn_update_Ax(kcrn_handle h, const int64_t *I_dev, const double *delta_dev, int m);

/* ---- LIBSVM / svmlight text -> CSR (host only; replaces sklearn.datasets.load_svmlight_file at
 * cubic_newton.py:52).  count: rows, stored entries and the smallest / largest feature index as written; parse:
 * fills labels[n_rows], indptr[n_rows + 1], indices[nnz], data[nnz] (indices as written: the zero- / one-based
 * decision and the per-row sort stay with the caller, as in sklearn).  threads < 1: all host threads.  The text
 * must be NUL-terminated at text[len]. */
int kcrn_svmlight_count(const char *text, int64_t len, int threads, int64_t *n_rows, int64_t *nnz,
                        int32_t *min_index, int32_t *max_index);
int kcrn_svmlight_parse(const char *text, int64_t len, int threads, int64_t n_rows, int64_t nnz, double *labels,
                        int64_t *indptr, int32_t *indices, double *data);

/* ---- introspection for the benchmark --------------------------------------------------------
 * counters[0..8): kernels launched, graph launches, HVPs, forward sweeps, transpose sweeps,
 * allreduces, bytes H2D, bytes D2H.  */
int kcrn_counters(kcrn_handle h, int64_t *counters8_host, int reset);
/* timing probe used by bench.py's roofline leg: about `reps` back-to-back launches of one sweep kernel (which = 0:
 * CSR pass of the HVP, 1: CSC pass of the HVP), launched the way the product launches it -- nodes of a captured
 * graph joined by programmatic edges, 20 per graph -- between two CUDA events on the handle's stream; returns the
 * mean duration per launch in microseconds.  which = 2: the two passes alternating (= plain HVPs, 10 per graph; still the
 * mean per launch). */
int kcrn_time_sweep(kcrn_handle h, int which, int reps, double *mean_us_host);
/* Benchmark utility (no reference counterpart): evict everything from L2 -- a buffer of `bytes` (choose > 126 MB) is
 * written on the handle's stream, a second one of the same size is read (the first one's dirty lines leave L2 as well),
 * both are freed; returns when that is done.  A write-only flush, or a flush buffer that stays allocated, slows an
 * L2-resident working set down for good (profiles/r3g_flush_probe.txt, r3h_flush_probe2.txt). */
int kcrn_flush_l2(kcrn_handle h, size_t bytes);
/* Host-only (no device needed): the tile list the sweeps would use for a matrix with this pointer array (rows of the
 * CSR copy / columns of the CSC copy) -- no reference counterpart; lets the CPU test suite check the tiling
 * invariants.  tiles4[i] = {first segment, first entry, z, entries | tag << 16}: z > 0 whole segments (tag 15: one
 * segment reduced by the whole CTA), z < 0 one piece of long segment -z-1; longtab4[l] = {first partial slot,
 * pieces, first entry of the segment, segment}.  Arrays are filled only when the capacities suffice; the counts
 * and the tile size (entries per thread, x 256 threads) are always returned. */
int kcrn_plan_tiles(const int32_t *indptr_host, int64_t n_segments, int32_t *tiles4_host, int64_t tiles_cap,
                    int64_t *n_tiles, int32_t *longtab4_host, int64_t long_cap, int64_t *n_long, int *items);
/* The balanced one-tile-per-SM plan of the persistent recurrence kernel (kcrn_persist.cu) for a matrix with this pointer
 * array on a device of `grid` SMs -- no reference counterpart; for the CPU test suite.  tiles12[i] = {first entry,
 * entries, first segment, segments, log2 threads per segment, head cut segment (-1: none), its piece, tail cut
 * segment, its piece, 0, 0, 0}; longtab4[l] = {first partial slot, pieces, segment, 0}.  *fits = 0: the matrix
 * needs more than `grid` tiles and the handle would keep the per-pass kernels. */
int kcrn_plan_persist(const int32_t *indptr_host, int64_t n_segments, int grid, int32_t *tiles12_host, int64_t tiles_cap,
                      int64_t *n_tiles, int32_t *longtab4_host, int64_t long_cap, int64_t *n_long, int *fits);
/* Which gather variant of the sweep kernel was measured fastest for this handle's matrices (no reference
 * counterpart; all variants compute the same bits): out4 = {CSR plain, CSR Lanczos, CSC plain, CSC Lanczos},
 * 0 = entries as stored, 1 = gather-ordered tiles, 2 = gather-ordered tiles with asynchronous (LDGSTS) gathers
 * (3 = experimental: values staged in shared memory; only on request).
 * The Lanczos entries are final after the first kcrn_lanczos / kcrn_krylov_trial / kcrn_krylov_run. */
int kcrn_gather_variants(kcrn_handle h, int *out4_host);

#ifdef __cplusplus
}
#endif
#endif /* KCRN_H */
