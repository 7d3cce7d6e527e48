#!/usr/bin/env python
"""Benchmark of the Krylov-CRN hot path (BASELINE.json: HVPs/s and outer iterations/s; HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload rcv1] [--m 10]

One "step" = one Krylov cubic-regularised-Newton outer iteration (gradient, m-step Lanczos = m HVPs, the
m x m sub-problem, trial point, value) on a synthetic sparse logistic-regression problem of the shape
BASELINE.json names (configs[1]: rcv1-shape, m = 10).  N > 1 (torchrun, one rank per GPU): every rank holds
one rcv1-shaped row shard of an N-times taller matrix (weak scaling) and each HVP / gradient / value adds
one NCCL allreduce; `value` then counts shard-HVPs (N per global HVP).

Printed JSON (rank 0, one line): value = HVPs/s with the iterate resident in HBM; e2e = the same through the
host-buffer API (numpy x in, numpy x out every step); roofline = the dominant sweep kernel timed live with
CUDA events against MEASURED_PEAKS.json; cpu_baseline = the CPU oracle port on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "krylov-cubic-regularized-newton_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC, UNIT = "krylov_crn_hvps_per_sec", "HVP/s"


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload, which):
    """DRAM bytes per launch of the CSR (0) / CSC (1) pass from the committed `ncu --set full` capture of this
    workload (profiles/ncu_traffic.json); None when there is no capture for it."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        if t.get("workload") != workload:
            return None
        return int(t["csc_pass_bytes_per_launch" if which else "csr_pass_bytes_per_launch"])
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        return False

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def make_shard(workload, rank):
    from kcrn_b200.synth import make_problem
    return make_problem(workload, seed=rank)


def cpu_oracle_run(P, m, steps, warmup, threads):
    """The CPU restatement of the reference path (oracle/), timed on the host cores.  Only used when the reference
    itself is not installed under baseline/_ref.  BLAS is held to one thread while it runs: the port's sweeps are
    OpenMP loops and a spinning OpenBLAS pool between them costs a factor of 20 (round-1 finding)."""
    from oracle import oracle as O
    from threadpoolctl import threadpool_limits
    with threadpool_limits(limits=1, user_api="blas"):
        L = O.OracleLogistic(P.indptr, P.indices, P.data, (P.n, P.d), P.b, threads=threads)
        opt = O.KrylovCRN(L, subspace_dim=m, tolerance=0)
        opt.x = np.array(P.x0, copy=True)
        L.reset()
        for _ in range(warmup):
            opt.step()
        h0 = L.calls["hvp"]
        t0 = time.perf_counter()
        for _ in range(steps):
            opt.step()
        dt = time.perf_counter() - t0
        hv = L.calls["hvp"] - h0
    return hv / dt, steps / dt, dt, hv


REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def reference_run(P, m, steps, warmup):
    """The UNMODIFIED reference (baseline/_ref/optimizer = the files of /root/reference/optimizer, installed by
    __graft_entry__.build()): Cubic_Krylov_LS(...).run(x0, it_max=warmup + steps) through its own public API, exactly
    as BASELINE.md section 4 prescribes -- matplotlib stubbed (opt_trace.py imports pyplot at module level), a counting
    wrapper on loss.hess_vec_prod, timing from the reference's own trace.ts (optimizer.py:129,151).  The first
    `warmup` iterations (they include the numba JIT of logsig) are dropped.  Runs in a child interpreter so that the
    reference's `optimizer` package and this repository's mirror of the same name never meet."""
    code = r"""
import json, os, sys, time, types
root, ref, m, steps, warmup, workload = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
mpl.pyplot = plt
sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt})
import importlib.util
spec = importlib.util.spec_from_file_location("kcrn_synth", os.path.join(root, "krylov-cubic-regularized-newton_b200", "kcrn_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec); sys.modules["kcrn_synth"] = synth; spec.loader.exec_module(synth)
sys.path.insert(0, ref)
import warnings; warnings.filterwarnings("ignore")
import numpy as np
import optimizer.cubic as RC
from optimizer.loss import LogisticRegression
assert os.path.realpath(RC.__file__).startswith(os.path.realpath(ref)), RC.__file__
P = synth.make_problem(workload, seed=0)
loss = LogisticRegression(P.scipy_csr(), P.b, l1=0, l2=0, store_mat_vec_prod=True)
opt = RC.Cubic_Krylov_LS(loss=loss, reg_coef=1e-3, subspace_dim=m, tolerance=0, label="reference", tqdm=False)
calls = {}
inner = loss.hess_vec_prod
def counted(x, v, *a, **k):
    calls[opt.it] = calls.get(opt.it, 0) + 1
    return inner(x, v, *a, **k)
loss.hess_vec_prod = counted
import io, contextlib
with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()):
    trace = opt.run(P.x0, it_max=warmup + steps)
its, ts = list(trace.its), list(trace.ts)
assert its[-1] == warmup + steps and warmup in its, (its[-5:], warmup)
dt = ts[-1] - ts[its.index(warmup)]
hv = sum(c for it, c in calls.items() if it >= warmup)
try:
    from threadpoolctl import threadpool_info
    pools = [{k: p.get(k) for k in ("user_api", "internal_api", "num_threads")} for p in threadpool_info()]
except Exception:
    pools = None
print(json.dumps({"hvps": hv, "seconds": dt, "steps": steps, "n": P.n, "d": P.d, "nnz": P.nnz, "pools": pools,
                  "f_last": float(opt.value), "cpu_count": os.cpu_count()}))
"""
    env = dict(os.environ)
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm must see the same thread pools at every N
    env.pop("OMP_NUM_THREADS", None)
    out = subprocess.run([sys.executable, "-c", code, ROOT, REF_DIR, str(m), str(steps), str(warmup), P],
                         capture_output=True, text=True, env=env, timeout=1500)
    if out.returncode != 0:
        raise RuntimeError("reference arm failed: " + out.stderr[-800:])
    return json.loads(out.stdout.strip().splitlines()[-1])


def cpu_reference(workload, m, steps, warmup):
    """(value HVP/s, outer it/s, seconds, HVPs, cpu_baseline dict) -- the reference when installed, else the port."""
    if os.path.isdir(os.path.join(REF_DIR, "optimizer")):
        r = reference_run(workload, m, steps, warmup)
        hps, ips = r["hvps"] / r["seconds"], r["steps"] / r["seconds"]
        blas = max([p["num_threads"] for p in (r["pools"] or []) if p.get("user_api") == "blas"] or [1])
        cb = {"value": hps, "unit": UNIT, "cores": 1, "kind": "reference", "outer_iters_per_sec": ips,
              "host_cpus": r["cpu_count"], "threadpools": r["pools"],
              "sample": f"{r['steps']} outer iterations ({r['hvps']} HVPs) of the same workload after {warmup} warm-up "
                        f"iterations, unmodified reference Cubic_Krylov_LS.run (baseline/_ref); scipy csr/csc matvec "
                        f"(~70% of its time) is single-threaded, BLAS pool {blas} threads for the dense vector ops: "
                        f"effective parallelism ~1 core"}
        return hps, ips, r["seconds"], r["hvps"], cb, (r["n"], r["d"], r["nnz"])
    P = make_shard(workload, 0)
    threads = os.cpu_count() or 1
    hps, ips, dt, hv = cpu_oracle_run(P, m, steps, warmup, threads)
    cb = {"value": hps, "unit": UNIT, "cores": threads, "kind": "port", "outer_iters_per_sec": ips,
          "sample": f"{steps} outer iterations ({hv} HVPs) of the same workload, CPU oracle port (oracle/, OpenMP sweeps, "
                    f"BLAS held to 1 thread); baseline/_ref is not installed on this box"}
    return hps, ips, dt, hv, cb, (P.n, P.d, P.nnz)


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 40)), max(1, min(args.warmup, 5))
    hps, ips, dt, hv, cb, (n, d, nnz) = cpu_reference(args.workload, args.m, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": hps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "outer_iters_per_sec": ips,
        "config": {"workload": f"{args.workload}-shape n={n} d={d} nnz={nnz}, Krylov CRN m={args.m}, x0=0.5, "
                               f"reg_coef=1e-3",
                   "note": "one shard's problem on the host CPU (the reference has no multi-process path); "
                           "iterations warmup..warmup+steps of Cubic_Krylov_LS.run from x0"},
        "cpu_baseline": cb,
        "e2e": {"value": hps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rcv1")
    ap.add_argument("--m", type=int, default=10)
    ap.add_argument("--reorth", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep-reps", type=int, default=200)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libkcrn has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from optimizer.cubic import Cubic_Krylov_LS
    from optimizer.loss import LogisticRegression

    W, K = max(3, args.warmup), max(1, args.steps)
    P = make_shard(args.workload, rank)
    loss = LogisticRegression(P.scipy_csr(), P.b, l1=0, l2=0, store_mat_vec_prod=True, n_global=P.n * world)
    eng = loss.engine
    if world > 1:
        from kcrn_b200.dist import join_communicator
        join_communicator(eng, rank, world)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed_steps(step_fn, count):
        """Sum of per-step CUDA-event times, L2 flushed (untimed) before every step; max over ranks."""
        evs = []
        barrier()
        for _ in range(count):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step_fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---------------- device-resident run: the real trajectory from x0 = 0.5 (cubic_newton.py:61)
    opt = Cubic_Krylov_LS(loss=loss, reg_coef=1e-3, subspace_dim=args.m, tolerance=0, reorth=bool(args.reorth),
                          label="bench", tqdm=False)
    opt.init_run(P.x0)
    for _ in range(W):
        opt.step()
    eng.counters(reset=True)
    n_hist = len(opt.history)
    clk = ClockSampler(local)
    clk.__enter__()            # sampled over every measured phase below (the device-resident run alone is ~10 ms)
    ms = timed_steps(opt.step, K)
    cnt = eng.counters()
    hvps = sum(h["hvps"] for h in opt.history[n_hist:])
    value = world * hvps / (ms * 1e-3)

    # ---------------- the same iterations with the step loop itself on the device (kcrn_krylov_run: one graph launch
    # for K outer iterations, line search in WHILE nodes); no flush between the iterations of one launch
    device_loop = None
    try:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # the same iterations as above: restart at x0, W warm-up iterations (they also capture the graph), then K
        v0 = loss.value(loss.to_device(P.x0))
        eng.krylov_run(max(W, K), args.m, bool(args.reorth), v0, 1e-3, 0.5, 0.1)     # sizes the buffers, captures
        loss.reset()
        v0 = loss.value(loss.to_device(P.x0))
        recs, _ = eng.krylov_run(W, args.m, bool(args.reorth), v0, 1e-3, 0.5, 0.1)
        state = (recs[-1][0], recs[-1][12], recs[-1][2])
        barrier()
        flush.fill_(1)
        ev0.record()
        recs, _ = eng.krylov_run(K, args.m, bool(args.reorth), *state[:2], 0.5, state[2])
        ev1.record()
        barrier()
        ms3 = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms3], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms3 = float(t.item())
        device_loop = {"ms_per_step": ms3 / len(recs), "outer_iters_per_sec": 1e3 * len(recs) / ms3,
                       "value": world * float(recs[:, 7].sum()) / (ms3 * 1e-3), "unit": UNIT, "steps": len(recs),
                       "note": "K outer iterations in ONE graph launch, line search on the device; one L2 flush "
                               "before the launch, none between its iterations"}
        loss.reset()
    except Exception as e:                                         # sharded without peer exchange: host loop only
        device_loop = {"unavailable": str(e)[:200]}

    # ---------------- end to end: numpy iterate in pinned host memory goes in and comes back every step
    opt2 = Cubic_Krylov_LS(loss=loss, reg_coef=1e-3, subspace_dim=args.m, tolerance=0, reorth=bool(args.reorth),
                           label="bench-e2e", tqdm=False)
    opt2.init_run(P.x0)
    x_host = torch.as_tensor(P.x0).clone().pin_memory()
    rec_host = torch.zeros(1, dtype=torch.float64).pin_memory()

    def e2e_step():
        opt2.x = x_host.to("cuda", non_blocking=True)          # H2D of the step's input
        opt2.step()
        x_host.copy_(opt2.x, non_blocking=True)                # D2H of the step's result
        rec_host[0] = opt2.value
        torch.cuda.current_stream().synchronize()

    for _ in range(W):
        e2e_step()
    n2 = len(opt2.history)
    ms2 = timed_steps(e2e_step, K)
    hvps2 = sum(h["hvps"] for h in opt2.history[n2:])
    e2e_value = world * hvps2 / (ms2 * 1e-3)

    # ---------------- roofline of the dominant kernel, timed live (back-to-back launches, CUDA events)
    peak, peak_src = measured_peak_gbs()
    torch.cuda.synchronize()
    us = [eng.time_sweep(w, args.sweep_reps) for w in (0, 1)]
    dom = int(us[1] >= us[0])
    gbs = [eng.sweep_bytes(w) / (us[w] * 1e-6) / 1e9 for w in (0, 1)]
    step_ms = ms / K
    hvp_us = 1e3 * step_ms / max(1.0, hvps / K)                 # whole step amortised per HVP
    # the captured m-step Lanczos graph alone (m HVPs + the vector kernels between them), back to back, at x0 (late
    # iterates break down early -- cubic.py:98 -- and the rest of the graph is no-op nodes)
    eng.set_x(loss.to_device(P.x0))
    eng.gradient(keep_internal=True)
    for _ in range(3):
        eng.lanczos(args.m, bool(args.reorth))
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    g0.record()
    for _ in range(reps):
        eng.lanczos(args.m, bool(args.reorth))
    g1.record()
    torch.cuda.synchronize()
    graph_us = 1e3 * g0.elapsed_time(g1) / reps
    graph_hvps = int(eng.lanczos_view()[3][3].item())
    lan_us = graph_us / max(1, graph_hvps)
    gather_variants = eng.gather_variants()       # which sweep-kernel variant the engine measured fastest per matrix
    clk.__exit__(None, None, None)
    clocks = clk.summary()
    roofline = {
        "bound": "hbm", "achieved": gbs[dom], "peak": peak, "unit": "GB/s", "frac": gbs[dom] / peak,
        "traffic": ncu_traffic(args.workload, dom),
        "kernel": "sweep_kernel<kColsScaled> (CSC pass: A' t)" if dom else "sweep_kernel<kRowsWeighted> (CSR pass: D.*(A v))",
        "kernel_us": us[dom], "algorithmic_bytes": eng.sweep_bytes(dom), "peak_source": peak_src,
        "other_pass": {"kernel_us": us[1 - dom], "achieved": gbs[1 - dom], "frac": gbs[1 - dom] / peak},
        "hvp": {"bytes": eng.hvp_bytes(), "us_two_sweeps": us[0] + us[1],
                "frac_two_sweeps": eng.hvp_bytes() / ((us[0] + us[1]) * 1e-6) / 1e9 / peak,
                "us_in_lanczos_graph": lan_us, "frac_in_lanczos_graph": eng.hvp_bytes() / (lan_us * 1e-6) / 1e9 / peak,
                "us_in_step": hvp_us, "frac_in_step": eng.hvp_bytes() / (hvp_us * 1e-6) / 1e9 / peak},
        "note": "CSR+CSC of this shape (36 MB) are L2-resident between sweeps, as in the real Lanczos loop; "
                "fractions are algorithmic bytes over the measured HBM copy peak",
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "outer_iters_per_sec": 1e3 * K / ms, "hvps_per_step": hvps / K,
        "config": {"workload": f"{args.workload}-shape n={P.n} d={P.d} nnz={P.nnz} per rank, Krylov CRN m={args.m}"
                               f"{' reorth' if args.reorth else ''}, x0=0.5, reg_coef=1e-3",
                   "sharding": "rows, one shard per rank, allreduce(d) per sweep" if world > 1 else "single GPU",
                   "l2": "256 MiB flush write (untimed) before every timed step; inside a step the matrix is "
                         "re-read 2m+2 times by design"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * P.d, "d2h_bytes_per_step": 8 * P.d + 120,
                "ms_per_step": ms2 / K, "outer_iters_per_sec": 1e3 * K / ms2},
        "gpu_launches": cnt["kernels"], "graph_launches": cnt["graph_launches"],
        "gather_variants": gather_variants,
        "roofline": roofline, "clocks": clocks, "device_loop": device_loop,
    }
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_reference(args.workload, args.m, min(K, 20), 3)[4]
    if rank == 0:
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
