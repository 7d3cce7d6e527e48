#!/usr/bin/env python
"""Secondary benchmark: BASELINE.json configs[3] -- full-Hessian CRN and SSCN (m = 10/50/100/500) on the
rcv1-shape problem, 1 B200.  Prints one JSON line; the headline contract lives in bench.py.

  dense A' D A (47 236^2, FP64 tensor cores): seconds, TFLOP/s over the upper-triangle flops actually issued
  SSCN: outer iterations / s per subspace dimension
  full-Hessian CRN: seconds per outer iteration (Hessian + cuSOLVER Cholesky solves of the dense secular equation)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "krylov-cubic-regularized-newton_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="rcv1")
    ap.add_argument("--sscn-iters", type=int, default=50)
    ap.add_argument("--crn-iters", type=int, default=1)
    ap.add_argument("--skip-crn", action="store_true")
    args = ap.parse_args()
    import torch

    from kcrn_b200.synth import make_problem
    from optimizer.cubic import SSCN, Cubic_LS
    from optimizer.loss import LogisticRegression

    P = make_problem(args.workload, seed=0)
    loss = LogisticRegression(P.scipy_csr(), P.b, l1=0, l2=0, store_mat_vec_prod=True)
    x = torch.as_tensor(P.x0, device="cuda")
    out = {"workload": f"{args.workload}-shape n={P.n} d={P.d} nnz={P.nnz}", "dtype": "f64"}
    from bench import ClockSampler
    clk = ClockSampler(0)
    clk.__enter__()

    # ---- dense Hessian
    H = torch.empty((P.d, P.d), dtype=torch.float64, device="cuda")
    loss._forward(x)
    loss.engine.hessian_dense(H)                       # warm-up (allocations, first launch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss.engine.hessian_dense(H)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3
    T = (P.d + 127) // 128
    n_pad = (P.n + 15) // 16 * 16
    flops = 2.0 * n_pad * 128 * 128 * (T * (T + 1) // 2)          # issued on the tensor cores (upper-triangle tiles)
    v = torch.as_tensor(np.random.default_rng(0).standard_normal(P.d), device="cuda")
    hv = loss.engine.hvp(v)
    err = float(((H @ v) - hv).norm() / hv.norm())
    out["dense_hessian"] = {"seconds": sec, "tflops_issued": flops / sec / 1e12, "includes": "memset + densify + DMMA GEMM",
                            "fp64_tensor_nominal_tflops": 40.0, "check_Hv_vs_hvp_rel_err": err,
                            "useful_dense_flops_2nd2": 2.0 * P.n * P.d * P.d}
    del H
    torch.cuda.empty_cache()

    # ---- SSCN
    out["sscn_iters_per_sec"] = {}
    for m in (10, 50, 100, 500):
        opt = SSCN(loss=loss, reg_coef=1e-3, subspace_dim=m, tolerance=1e-9, label="s", tqdm=False)
        opt.rng = np.random.default_rng(42)
        opt.init_run(P.x0)
        for _ in range(3):
            opt.step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.sscn_iters):
            opt.step()
        torch.cuda.synchronize()
        out["sscn_iters_per_sec"][str(m)] = args.sscn_iters / (time.perf_counter() - t0)

    # ---- full-Hessian CRN (dense 47k^2 Cholesky per Newton evaluation: seconds per outer iteration)
    if not args.skip_crn:
        opt = Cubic_LS(loss=loss, reg_coef=1e-3, cubic_solver="full", tolerance=1e-8, label="c", tqdm=False)
        opt.init_run(P.x0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.crn_iters):
            opt.step()
        torch.cuda.synchronize()
        out["full_crn"] = {"seconds_per_outer_iteration": (time.perf_counter() - t0) / args.crn_iters,
                           "newton_iterations": opt.solver_it, "f": opt.value}
    clk.__exit__(None, None, None)
    out["clocks"] = clk.summary()
    out["reference_cpu_sscn_iters_per_sec"] = {"10": 165, "50": 128, "100": 84, "500": 11.0,
                                               "source": "BASELINE.md section 2 (unmodified reference, survey container, 8 cores)"}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
