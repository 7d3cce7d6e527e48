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
    """DRAM bytes per launch of the CSR (0) / CSC (1) pass, or of the rows-Lanczos + columns-Lanczos pair (2), from the
    committed `ncu --set full` capture of this workload (profiles/ncu_traffic.json); None when there is no capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        if t.get("workload") != workload:
            return None
        return int(t[("csr_pass_bytes_per_launch", "csc_pass_bytes_per_launch", "lanczos_pair_bytes_per_hvp")[which]])
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


NVLINK_GBS = 770.0      # measured peer copy per direction per GPU on this pool (B200_PROFILING.md); nominal 900


def strong_case(tag, n, d, per_row, m, steps, rank, world, local, peak):
    """One strong-scaled configuration of BASELINE.json (configs[2] news20-shape, configs[4] 1 B non-zeros): the SAME
    global problem at every N -- n rows split evenly over the ranks, every rank generates its own rows on its device
    (one shared column-popularity permutation) -- Krylov-CRN outer iterations with one allreduce of the d-vector per
    transpose sweep.  Timed with CUDA events over `steps` outer iterations after one warm-up iteration, max over
    ranks; clocks sampled during the timed region.  The step loop is what Cubic_Krylov_LS.step does
    (optimizer/cubic.py: M = reg_coef * beta first, trial, M / beta on rejection)."""
    import torch
    import torch.distributed as dist
    from kcrn_b200.dist import join_communicator
    from kcrn_b200.engine import Engine
    from kcrn_b200.synth import make_problem_device

    n_loc = n // world
    indptr, indices, data, b = make_problem_device(n_loc, d, float(per_row), seed=1000 + rank, perm_seed=7)
    nnz = int(indices.numel())
    eng = Engine(indptr, indices, data, b, (n_loc, d), n_global=n_loc * world)
    del indptr, indices, data, b
    torch.cuda.empty_cache()
    if world > 1:
        join_communicator(eng, rank, world)
    x = torch.full((d,), 0.5, device="cuda", dtype=torch.float64)
    state = {"x": x, "M": 1e-3, "r0": 0.1, "value": None, "hvps": 0}

    def step():
        eng.set_x(state["x"]) if state["value"] is None else None
        if state["value"] is None:
            state["value"] = eng.value()
        M, tries, first = state["M"] * 0.5, 0, True
        while True:
            x_new = eng.new_vector()
            val, dec, r, its, status, m_eff, gnorm, hv, dx = eng.krylov_trial(first, m, False, M, state["r0"], 1e-8, 100,
                                                                               x_out=x_new)
            if first:
                state["hvps"] += hv
            first = False
            if not (val > state["value"] - dec and tries < 20):
                break
            M, tries = M / 0.5, tries + 1
        state.update(x=x_new, M=M, r0=r, value=val)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    clk = ClockSampler(local)                  # started ahead of the warm-up step: nvidia-smi needs ~0.1 s to deliver its
    clk.__enter__()                            # first sample and the timed region of the small case is shorter than that
    step()                                     # warm-up: captures the graphs, measures the gather variants
    sync()
    x_first = state["x"].clone()               # second iterate: where the K6 timing below runs its recurrences
    h0 = state["hvps"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    sync()
    clk.__exit__(None, None, None)
    t = torch.tensor([e0.elapsed_time(e1), float(nnz)], dtype=torch.float64, device="cuda")
    if world > 1:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms, nnz_total = float(tm[0].item()), int(t[1].item())
    else:
        ms, nnz_total = float(t[0].item()), nnz
    hv = state["hvps"] - h0
    us_hvp = 1e3 * ms / max(1, hv)             # whole step amortised per HVP (gradient, sub-problem, value included)
    hbm_bytes = 24 * nnz + 28 * n_loc + 28 * d + 8                       # this rank's two sweeps
    ar_bytes = 2.0 * (world - 1) / world * 8 * d if world > 1 else 0.0    # ring allreduce of the d-vector, per rank
    roof_us = hbm_bytes / (peak * 1e9) * 1e6 + ar_bytes / (NVLINK_GBS * 1e9) * 1e6
    reorth = None
    if world == 1 and d > 1000000 and nnz < 50000000:
        # north_star K6 where it is bandwidth-bound (d = 1.36 M: V[:j+1] is j x 10.8 MB): the recurrence with full
        # re-orthogonalisation minus the one without, against three reads of V[:j+1] per step
        def rec_ms(ro):
            eng.lanczos(m, ro)
            torch.cuda.synchronize()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(2):
                eng.lanczos(m, ro)
            b_.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b_) / 2, int(eng.lanczos_view()[3][3].item())
        # at the iterate the warm-up step started from (kept in x_first): later iterates -- and x0 = 0.5 itself -- break the
        # recurrence down early (cubic.py:98) and the rest of the graph is no-op nodes; the bytes follow the steps taken
        eng.set_x(x_first)
        eng.gradient(keep_internal=True)
        (t_no, _), (t_ro, steps_ro) = rec_ms(False), rec_ms(True)
        vbytes = 3 * 8 * d * sum(j + 1 for j in range(max(0, steps_ro - 1)))
        reorth = {"recurrence_ms": t_ro, "recurrence_ms_without": t_no, "hvps": steps_ro, "basis_bytes": vbytes,
                  "achieved_gbs": vbytes / ((t_ro - t_no) * 1e-3) / 1e9, "frac": vbytes / ((t_ro - t_no) * 1e-3) / 1e9 / peak}
    out = {"workload": f"{tag}: n={n_loc * world} d={d} nnz={nnz_total}, Krylov CRN m={m}, rows split evenly over {world} GPU(s)",
           "reorth": reorth,
           "scaling": "strong", "hvps_per_sec": 1e6 / us_hvp, "outer_iters_per_sec": 1e3 * steps / ms, "steps": steps,
           "us_per_hvp_in_step": us_hvp, "hvps_per_step": hv / steps,
           "roofline": {"hbm_bytes_per_rank": hbm_bytes, "allreduce_bytes_per_rank": ar_bytes, "roofline_us": roof_us,
                        "frac": roof_us / us_hvp, "peak_hbm_gbs": peak, "nvlink_gbs": NVLINK_GBS},
           "exchange": "NCCL allreduce" if world > 1 else "none", "clocks": clk.summary()}
    eng.close()
    del eng
    torch.cuda.empty_cache()
    return out


def sharded_parity(loss, P, rank, world, workload):
    """Row-sharded HVP / gradient / value against scipy on the STACKED matrix (rank 0 rebuilds every shard), and the
    replicas against each other bit for bit."""
    import scipy.sparse as sp
    import scipy.special
    import torch
    import torch.distributed as dist

    rng = np.random.default_rng(123)
    x = 0.5 * np.ones(P.d) + 0.01 * rng.standard_normal(P.d)
    v = rng.standard_normal(P.d)
    xd, vd = loss.to_device(x), loss.to_device(v)
    hv = loss.hess_vec_prod(xd, vd)
    g = loss.gradient(xd)
    f = torch.tensor([loss.value(xd)], dtype=torch.float64, device="cuda")
    pack = torch.cat([hv, g, f])
    every = [torch.empty_like(pack) for _ in range(world)]
    dist.all_gather(every, pack)
    if rank != 0:
        return None
    same = all(torch.equal(every[0], e) for e in every[1:])
    shards = [make_shard(workload, r) for r in range(world)]
    A = sp.vstack([s.scipy_csr() for s in shards]).tocsr()
    b = (np.concatenate([s.b for s in shards]) + 1) / 2
    n = A.shape[0]
    Ax = A @ x
    a = scipy.special.expit(Ax)
    hv_ref = A.T @ (a * (1 - a) * (A @ v)) / n
    g_ref = A.T @ (a - b) / n
    f_ref = float(np.mean((1 - b) * Ax - np.where(Ax < -33, Ax, -np.log1p(np.exp(-np.clip(Ax, -700, 700))))))
    got = every[0].cpu().numpy()
    d = P.d
    rel = lambda u, w: float(np.linalg.norm(u - w) / np.linalg.norm(w))
    return {"replicas_bit_equal": bool(same), "hvp_rel_err": rel(got[:d], hv_ref), "grad_rel_err": rel(got[d:2 * d], g_ref),
            "value_rel_err": abs(got[2 * d] - f_ref) / abs(f_ref), "against": "scipy on the stacked matrix of all shards"}


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
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaled C3 / C5 cases")
    ap.add_argument("--c5-rows", type=float, default=10e6)
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

    # L2 flush between timed steps (untimed, kcrn_flush_l2): 256 MiB (> 126 MB L2) are written, another 256 MiB are read,
    # both buffers are freed.  The read pushes the dirty lines of the write out again: after a write-only flush -- or with a
    # flush buffer that stays allocated -- an L2-resident working set stays slower for good (recurrence at this shape 20.5 ->
    # 21.6 / 22.5 us per HVP; profiles/r3g_flush_probe.txt, r3h_flush_probe2.txt, r3g_bench_vs_microbench.txt): a property of
    # the harness, not of the library, which in production shares the GPU with nothing.  The effect is reported
    # (`roofline.resident_flush_buffer`).  No idle time is added behind the flush: a host sleep of 2 ms (KCRN_BENCH_FLUSH_SLEEP_MS)
    # makes the step behind it slower, not faster (0.327 -> 0.342 ms; with the line search on the device 0.36 -> 0.39:
    # a graph launched into an idle GPU starts later), and the lead time of the nvidia-smi sampler does not matter
    # (profiles/r4c_bench_harness_ab.txt; r4a / r4b_flush_step_probe.txt for the step behind each flush procedure).
    FLUSH_BYTES = 256 << 20
    FLUSH_SLEEP_MS = float(os.environ.get("KCRN_BENCH_FLUSH_SLEEP_MS", "0"))

    def flush_l2():
        eng.flush_l2(FLUSH_BYTES)
        if FLUSH_SLEEP_MS > 0:
            time.sleep(1e-3 * FLUSH_SLEEP_MS)

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
            flush_l2()
            if world > 1:
                barrier()          # the flush takes 1-2 ms and a different time on every rank: without this the ranks enter
                                   # the step that far apart and the timed region of the early one is the other's flush
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
    DEVICE_STEP = os.environ.get("KCRN_BENCH_DEVICE_STEP", "0") != "0"
    opt = Cubic_Krylov_LS(loss=loss, reg_coef=1e-3, subspace_dim=args.m, tolerance=0, reorth=bool(args.reorth),
                          label="bench", tqdm=False, device_step=DEVICE_STEP)
    opt.init_run(P.x0)
    for _ in range(W):
        opt.step()
    eng.counters(reset=True)
    n_hist = len(opt.history)
    clk = ClockSampler(local)
    clk.__enter__()            # sampled over every measured phase below (the device-resident run alone is ~10 ms)
    time.sleep(float(os.environ.get("KCRN_BENCH_SAMPLER_LEAD_S", "0")))
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
        flush_l2()
        if world > 1:
            barrier()              # as in timed_steps: the ranks leave the flush at different times
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
                           label="bench-e2e", tqdm=False, device_step=DEVICE_STEP)
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

    # ---------------- roofline, timed live with CUDA events.  The kernels that own ~3/4 of a step are the two Lanczos
    # instantiations of the sweep kernel (rows-Lanczos + columns-Lanczos = one HVP inside the captured recurrence), so
    # THEY are the roofline entry; the plain passes (launched back to back as graph nodes) are kept as `other`.
    peak, peak_src = measured_peak_gbs()
    torch.cuda.synchronize()
    us = [eng.time_sweep(w, args.sweep_reps) for w in (0, 1)]
    gbs = [eng.sweep_bytes(w) / (us[w] * 1e-6) / 1e9 for w in (0, 1)]
    step_ms = ms / K
    hvp_us = 1e3 * step_ms / max(1.0, hvps / K)                 # whole step amortised per HVP
    # the captured m-step recurrence alone (m HVPs + the vector kernels between them), back to back, at x0 (late
    # iterates break down early -- cubic.py:98 -- and the rest of the graph is no-op nodes)
    eng.set_x(loss.to_device(P.x0))
    eng.gradient(keep_internal=True)

    def time_recurrence(m, reorth, reps=20):
        for _ in range(3):
            eng.lanczos(m, reorth)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(reps):
            eng.lanczos(m, reorth)
        g1.record()
        torch.cuda.synchronize()
        return 1e3 * g0.elapsed_time(g1) / reps, int(eng.lanczos_view()[3][3].item())

    graph_us, graph_hvps = time_recurrence(args.m, bool(args.reorth))
    lan_us = graph_us / max(1, graph_hvps)
    # the same recurrence while a 256 MiB buffer that has been written once stays allocated (what a persistent flush buffer
    # did to every number of this bench until the last session of round 2)
    resident = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device="cuda")
    resident.fill_(1)
    graph_us_res, hv_res = time_recurrence(args.m, bool(args.reorth))
    del resident
    torch.cuda.empty_cache()
    resident_info = {"us_in_lanczos_graph": graph_us_res / max(1, hv_res),
                     "note": "a 256 MiB buffer written once and left allocated; freed again before anything else is timed"}
    # full re-orthogonalisation (north_star K6; lanczos_update_kernel stages 1-3): cost of the recurrence with it minus
    # without, against the bytes of V it reads (three passes over V[:j+1] per step as implemented)
    reorth_info = None
    if world == 1:
        reorth_info = {}
        for mm in sorted({args.m, 50}):
            t0, h0 = time_recurrence(mm, False, reps=10)
            t1, h1 = time_recurrence(mm, True, reps=10)
            vbytes = 3 * 8 * P.d * sum(j + 1 for j in range(mm - 1))
            extra = max(t1 - t0, 1e-3)
            reorth_info[f"m{mm}"] = {"recurrence_us": t1, "recurrence_us_without": t0, "hvps": h1,
                                     "reorth_us_per_step": extra / max(1, mm - 1), "basis_bytes": vbytes,
                                     "achieved_gbs": vbytes / (extra * 1e-6) / 1e9, "frac": vbytes / (extra * 1e-6) / 1e9 / peak}
        eng.lanczos(args.m, bool(args.reorth))
    gather_variants = eng.gather_variants()       # which sweep-kernel variant the engine measured fastest per matrix
    clk.__exit__(None, None, None)
    clocks = clk.summary()
    dom = int(us[1] >= us[0])
    roofline = {
        "bound": "hbm", "achieved": eng.hvp_bytes() / (lan_us * 1e-6) / 1e9, "peak": peak, "unit": "GB/s",
        "frac": eng.hvp_bytes() / (lan_us * 1e-6) / 1e9 / peak,
        "traffic": ncu_traffic(args.workload, 2),
        "kernel": "sweep_kernel<kRowsLanczos> + sweep_kernel<kColsLanczos>: one HVP inside the captured Lanczos recurrence "
                  "(the pair is timed: CUDA events around the graph, divided by its HVPs)",
        "kernel_us": lan_us, "algorithmic_bytes": eng.hvp_bytes(), "peak_source": peak_src,
        "other": {"kernel": "plain passes kRowsWeighted / kColsScaled, back to back as graph nodes",
                  "kernel_us": us, "achieved": gbs, "frac": [g / peak for g in gbs],
                  "traffic": [ncu_traffic(args.workload, 0), ncu_traffic(args.workload, 1)]},
        "hvp": {"bytes": eng.hvp_bytes(), "us_two_sweeps": us[0] + us[1],
                "frac_two_sweeps": eng.hvp_bytes() / ((us[0] + us[1]) * 1e-6) / 1e9 / peak,
                "us_in_lanczos_graph": lan_us, "frac_in_lanczos_graph": eng.hvp_bytes() / (lan_us * 1e-6) / 1e9 / peak,
                "us_in_step": hvp_us, "frac_in_step": eng.hvp_bytes() / (hvp_us * 1e-6) / 1e9 / peak},
        "reorth": reorth_info,
        "resident_flush_buffer": resident_info,
        "note": "CSR+CSC of this shape (36 MB) are L2-resident between sweeps, as in the real Lanczos loop; "
                "fractions are algorithmic bytes over the measured HBM copy peak",
    }
    parity = sharded_parity(loss, P, rank, world, args.workload) if world > 1 else None

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "outer_iters_per_sec": 1e3 * K / ms, "hvps_per_step": hvps / K,
        "config": {"workload": f"{args.workload}-shape n={P.n} d={P.d} nnz={P.nnz}{' per rank' if world > 1 else ''}, Krylov CRN m={args.m}"
                               f"{' reorth' if args.reorth else ''}, x0=0.5, reg_coef=1e-3",
                   "sharding": "rows, one shard per rank, allreduce(d) per sweep" if world > 1 else "single GPU",
                   "l2": "before every timed step (untimed, kcrn_flush_l2) 256 MiB are written, another 256 MiB read, both "
                         "freed: the write evicts the matrices, the read evicts the write's dirty lines, which otherwise "
                         "linger and slow the L2-resident step down; inside a step the matrix is re-read 2m+2 times by design",
                   "step": "Cubic_Krylov_LS.step(): one graph launch + one 96-byte record per trial of the line search "
                           "(kcrn_krylov_trial); `device_loop` = K iterations in one launch (kcrn_krylov_run)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * P.d, "d2h_bytes_per_step": 8 * P.d + 120,
                "ms_per_step": ms2 / K, "outer_iters_per_sec": 1e3 * K / ms2},
        "gpu_launches": cnt["kernels"], "graph_launches": cnt["graph_launches"],
        "gather_variants": gather_variants,
        "roofline": roofline, "clocks": clocks, "device_loop": device_loop,
    }
    if parity is not None:
        line["parity"] = parity
    eng.close()
    del opt, opt2, loss, eng
    torch.cuda.empty_cache()
    if not args.no_strong:
        # the strong-scaled configurations the north star names, the same global problem at every N
        strong = {}
        for tag, n_, d_, per_row, m_, steps_ in (("c3 news20-shape", 19996, 1355191, 450.0, 50, 20),
                                                  ("c5 1e9 non-zeros", int(args.c5_rows), 1000000, 116.0, 100, 2)):
            try:
                strong[tag.split()[0]] = strong_case(tag, n_, d_, per_row, m_, steps_, rank, world, local, peak)
            except Exception as e:
                strong[tag.split()[0]] = {"unavailable": f"{type(e).__name__}: {str(e)[:300]}"}
        line["strong"] = strong
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_reference(args.workload, args.m, min(K, 20), 3)[4]
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
