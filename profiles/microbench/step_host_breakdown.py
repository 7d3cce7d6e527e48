"""Where does the host-driven step (Cubic_Krylov_LS.step, per-trial path) spend its wall time?  The GPU idles while the
host works and the host waits while the GPU works, so the segments add up: Python before the C call, the C call
(kcrn_krylov_trial: parameter copy, graph launch, record copy, synchronise), Python behind it; GPU time of the graph from
CUDA events recorded around the C call."""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import torch
import bench
torch.cuda.set_device(0)
from optimizer.cubic import Cubic_Krylov_LS
from optimizer.loss import LogisticRegression

P = bench.make_shard("rcv1", 0)
loss = LogisticRegression(P.scipy_csr(), P.b, l1=0, l2=0, store_mat_vec_prod=True, n_global=P.n)
eng = loss.engine
T = {}


def timed(obj, name, key):
    f = getattr(obj, name)

    def g(*a, **k):
        t0 = time.perf_counter()
        r = f(*a, **k)
        T[key] = T.get(key, 0.0) + time.perf_counter() - t0
        T[key + "_n"] = T.get(key + "_n", 0) + 1
        return r
    setattr(obj, name, g)


timed(eng, "krylov_trial", "c_call")
timed(eng, "new_vector", "new_vector")
timed(loss, "_note_value", "note_value(clone)")
timed(loss, "adopt_iterate", "adopt")
timed(loss, "_forward", "forward_check")
timed(eng, "_use_stream", "use_stream")

opt = Cubic_Krylov_LS(loss=loss, reg_coef=1e-3, subspace_dim=10, tolerance=0, label="b", tqdm=False, device_step=False)
opt.init_run(P.x0)
for _ in range(5):
    opt.step()
for rounds in range(2):
    T.clear()
    N = 40
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(N):
        opt.step()
    tot = time.perf_counter() - t0
    trials = T["c_call_n"]
    print(f"{N} steps, {trials} trials: {1e6 * tot / N:.1f} us per step, {1e6 * tot / trials:.1f} us per trial")
    for k in sorted(k for k in T if not k.endswith("_n")):
        print(f"   {k:22s} {1e6 * T[k] / N:8.1f} us per step   ({T[k + '_n']} calls, {1e6 * T[k] / T[k + '_n']:.1f} us each)")
    print(f"   python outside these     {1e6 * (tot - T['c_call'] - T['new_vector'] - T['note_value(clone)'] - T['adopt'] - T['forward_check']) / N:8.1f} us per step")

# GPU time of one trial graph: events around the C call (the call ends with a synchronise)
opt.init_run(P.x0); opt.value = None
for _ in range(5):
    opt.step()
ev = []
orig = eng.krylov_trial


def trial_ev(*a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = orig(*a, **k)
    e1.record()
    ev.append((bool(a[0]), e0, e1))
    return r


eng.krylov_trial = trial_ev
for _ in range(40):
    opt.step()
torch.cuda.synchronize()
for first in (True, False):
    ts = sorted(a.elapsed_time(b) for f, a, b in ev if f == first)
    if ts:
        print(f"trial graph (first={first}): {len(ts)} launches, events around the C call: mean {1e3 * sum(ts) / len(ts):.1f} us, median {1e3 * ts[len(ts) // 2]:.1f} us")
