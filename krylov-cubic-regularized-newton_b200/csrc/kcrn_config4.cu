// Config-4 rows (full-Hessian CRN and SSCN): dense A' D A, column-block gradient / Hessian and the
// incremental A x update (optimizer/loss.py:234-264, :279-281).
#include "kcrn_internal.cuh"

namespace kcrn {

int config4_hessian_dense(kcrn_problem *, double *) {
    set_error("kcrn_hessian_dense: not built yet");
    return KCRN_E_ARG;
}
int config4_partial_gradient(kcrn_problem *, const int64_t *, int, double *) {
    set_error("kcrn_partial_gradient: not built yet");
    return KCRN_E_ARG;
}
int config4_partial_hessian(kcrn_problem *, const int64_t *, int, double *) {
    set_error("kcrn_partial_hessian: not built yet");
    return KCRN_E_ARG;
}
int config4_update_Ax(kcrn_problem *, const int64_t *, const double *, int) {
    set_error("kcrn_update_Ax: not built yet");
    return KCRN_E_ARG;
}

}  // namespace kcrn
