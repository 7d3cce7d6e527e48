// Device-resident outer loop of Cubic_Krylov_LS: the bookkeeping the reference does
// in Python between the kernels of one outer iteration -- "try a smaller M first", the acceptance test of the line
// search, M / beta on rejection, the commit of (x, value, M, r0), the stopping rule |x_new - x| < tolerance and the
// best-point tracking of Oracle.value -- as one-thread kernels that drive CUDA graph WHILE nodes.
// Reference: optimizer/cubic.py:265-309 (step), optimizer/optimizer.py:100-113 (check_convergence),
// optimizer/loss.py:66-73 (f_opt / x_opt).
#include "kcrn_internal.cuh"
#include "kcrn_device.cuh"

namespace kcrn {

namespace {

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// host parameters (already copied into ctl by a memcpy node in front of the graph) -> running state
__global__ void run_init_kernel(RunCtl *ctl) {
    if (threadIdx.x == 0) {
        ctl->iter = 0; ctl->tries = 0; ctl->status = 0; ctl->improved = 0; ctl->copy_best = 0; ctl->solver_it = 0;
        ctl->max_iters = (int)ctl->max_iters_d; ctl->max_tries = (int)ctl->max_tries_d; ctl->M = 0.0;
        ctl->t0 = global_ns();
    }
}

// top of an outer iteration: M = reg_coef * beta (cubic.py:286), parameters of the first sub-problem solve, and
// the line-search loop armed
__global__ void run_begin_kernel(RunCtl *ctl, double *step_params, cudaGraphConditionalHandle trial_loop) {
    if (threadIdx.x == 0) {
        ctl->M = ctl->reg_coef * ctl->beta;
        ctl->tries = 0;
        step_params[0] = ctl->M;
        step_params[1] = ctl->r0;
        cudaGraphSetConditional(trial_loop, 1);
    }
}

// after a trial (record at rec[0..12), see record_kernel): best point seen so far (loss.py:70-72), then the
// acceptance test `value_new > value - model_decrease and tries < max_tries` (cubic.py:297): reject -> M / beta and
// another trip through the loop body; accept -> leave the loop
__global__ void run_decide_kernel(RunCtl *ctl, const double *rec, double *step_params,
                                  cudaGraphConditionalHandle trial_loop) {
    if (threadIdx.x == 0) {
        const double value_new = rec[0], decrease = rec[1];
        const int status = (int)rec[4];
        ctl->copy_best = 0;
        if (status != 0) {                       // T + r I not positive definite: stop everything, the host raises
            ctl->status = status;
            cudaGraphSetConditional(trial_loop, 0);
            return;
        }
        if (value_new < ctl->f_opt) { ctl->f_opt = value_new; ctl->copy_best = 1; ctl->improved = 1; }
        if (value_new > ctl->value - decrease && ctl->tries < ctl->max_tries) {
            ctl->M = ctl->M / ctl->beta;
            ctl->tries += 1;
            step_params[0] = ctl->M;
            cudaGraphSetConditional(trial_loop, 1);
        } else {
            cudaGraphSetConditional(trial_loop, 0);
        }
    }
}

// x_opt = trial point when the decide kernel saw a new best value
__global__ void __launch_bounds__(kVecThreads) run_best_kernel(const RunCtl *ctl, const double *__restrict__ x,
                                                               double *__restrict__ x_opt, int64_t d) {
    if (!ctl->copy_best) return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d; i += (int64_t)gridDim.x * blockDim.x)
        x_opt[i] = x[i];
}

// xs[iter] = accepted iterate (the trace of optimizer.py:155-157), when the caller keeps iterates
__global__ void __launch_bounds__(kVecThreads) run_snapshot_kernel(const RunCtl *ctl, const double *__restrict__ x,
                                                                   double *__restrict__ xs, int64_t d) {
    if (ctl->status != 0) return;
    double *dst = xs + (size_t)ctl->iter * (size_t)d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = x[i];
}

// end of an outer iteration: commit (x is already in place), the 16-double record of the iteration, and the
// stopping rules of the step loop
__global__ void run_commit_kernel(RunCtl *ctl, const double *rec, double *recs, cudaGraphConditionalHandle outer_loop) {
    if (threadIdx.x == 0) {
        if (ctl->status != 0) { cudaGraphSetConditional(outer_loop, 0); return; }
        ctl->value = rec[0];
        ctl->reg_coef = ctl->M;
        ctl->r0 = rec[2];
        ctl->solver_it += (long long)rec[3];
        double *out = recs + (size_t)ctl->iter * 16;
        for (int i = 0; i < 11; ++i) out[i] = rec[i];
        out[11] = ctl->f_opt;
        out[12] = ctl->M;
        out[13] = (double)ctl->tries;
        out[14] = (double)(global_ns() - ctl->t0) * 1e-9;
        out[15] = (double)ctl->improved;
        ctl->iter += 1;
        const bool small_step = ctl->tol > 0.0 && rec[8] < ctl->tol;          // optimizer.py:110
        cudaGraphSetConditional(outer_loop, (ctl->iter < ctl->max_iters && !small_step) ? 1 : 0);
    }
}

}  // namespace

int launch_run_init(kcrn_problem *h) {
    run_init_kernel<<<1, 32, 0, h->stream>>>(h->run_ctl);
    KCRN_LAUNCH_CHECK();
    h->counters[0] += 1;
    return KCRN_OK;
}

int launch_run_begin(kcrn_problem *h, unsigned long long trial_loop) {
    run_begin_kernel<<<1, 32, 0, h->stream>>>(h->run_ctl, h->step_params, (cudaGraphConditionalHandle)trial_loop);
    KCRN_LAUNCH_CHECK();
    h->counters[0] += 1;
    return KCRN_OK;
}

int launch_run_decide(kcrn_problem *h, unsigned long long trial_loop) {
    run_decide_kernel<<<1, 32, 0, h->stream>>>(h->run_ctl, h->scal + 4, h->step_params, (cudaGraphConditionalHandle)trial_loop);
    KCRN_LAUNCH_CHECK();
    int64_t b = (h->d + kVecThreads - 1) / kVecThreads;
    if (b > kNumSm * 8) b = kNumSm * 8;
    run_best_kernel<<<(int)b, kVecThreads, 0, h->stream>>>(h->run_ctl, h->x_cur, h->run_x_opt, h->d);
    KCRN_LAUNCH_CHECK();
    h->counters[0] += 2;
    return KCRN_OK;
}

int launch_run_commit(kcrn_problem *h, unsigned long long outer_loop) {
    if (h->run_xs) {
        int64_t b = (h->d + kVecThreads - 1) / kVecThreads;
        if (b > kNumSm * 8) b = kNumSm * 8;
        run_snapshot_kernel<<<(int)b, kVecThreads, 0, h->stream>>>(h->run_ctl, h->x_cur, h->run_xs, h->d);
        KCRN_LAUNCH_CHECK();
        h->counters[0] += 1;
    }
    run_commit_kernel<<<1, 32, 0, h->stream>>>(h->run_ctl, h->scal + 4, h->run_recs, (cudaGraphConditionalHandle)outer_loop);
    KCRN_LAUNCH_CHECK();
    h->counters[0] += 1;
    return KCRN_OK;
}

}  // namespace kcrn
