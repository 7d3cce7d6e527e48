"""Why does the host-driven step cost 0.41 ms behind kcrn_flush_l2 (write + read + free) and 0.32 ms behind a write-only
flush kernel?  Per flush procedure: CUDA-event time and host wall time of the step that follows, and of the two steps
after that one (no flush in between)."""
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
resident = None


def spin(us):
    torch.cuda._sleep(int(us * 1965))          # cycles at the SM clock of this part


def f_none():
    torch.cuda.synchronize()


def f_async_write():
    global resident
    if resident is None:
        resident = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    resident.fill_(1)


def f_kcrn():
    eng.flush_l2(256 << 20)


def f_kcrn_sleep2():
    eng.flush_l2(256 << 20)
    time.sleep(0.002)


def f_kcrn_sleep20():
    eng.flush_l2(256 << 20)
    time.sleep(0.02)


def f_kcrn_spin():
    eng.flush_l2(256 << 20)
    spin(300)


def f_kcrn_sleep_spin():
    eng.flush_l2(256 << 20)
    time.sleep(0.002)
    spin(300)


def f_none_spin():
    torch.cuda.synchronize()
    spin(300)


def run(tag, flush, count=20, follow=2):
    opt = Cubic_Krylov_LS(loss=loss, reg_coef=1e-3, subspace_dim=10, tolerance=0, label="b", tqdm=False)
    opt.init_run(P.x0)
    for _ in range(5):
        opt.step()
    evs, host = [[] for _ in range(follow + 1)], [[] for _ in range(follow + 1)]
    torch.cuda.synchronize()
    for _ in range(count // (follow + 1) + 1):
        flush()
        for k in range(follow + 1):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            t0 = time.perf_counter()
            opt.step()
            t1 = time.perf_counter()
            b.record()
            evs[k].append((a, b)); host[k].append(1e3 * (t1 - t0))
    torch.cuda.synchronize()
    out = []
    for k in range(follow + 1):
        e = sorted(a.elapsed_time(b) for a, b in evs[k]); h = sorted(host[k])
        out.append(f"step+{k}: gpu {sum(e) / len(e):.3f} (med {e[len(e) // 2]:.3f}) host {sum(h) / len(h):.3f}")
    if resident is not None and flush is not f_async_write:
        pass
    print(f"{tag:58s} " + " | ".join(out), flush=True)


run("no flush, synchronize only", f_none)
run("no flush, synchronize + 300 us spin kernel", f_none_spin)
run("kcrn_flush_l2 (write, read, free both)", f_kcrn)
run("kcrn_flush_l2 + host sleep 2 ms", f_kcrn_sleep2)
run("kcrn_flush_l2 + host sleep 20 ms", f_kcrn_sleep20)
run("kcrn_flush_l2 + 300 us spin kernel", f_kcrn_spin)
run("kcrn_flush_l2 + sleep 2 ms + 300 us spin kernel", f_kcrn_sleep_spin)
run("no flush, synchronize only (again)", f_none)
run("resident buffer written asynchronously (round-1 flush)", f_async_write)
resident = None
torch.cuda.empty_cache()
run("no flush, synchronize only (buffer freed)", f_none)

# the device loop one iteration per launch, behind the same flush: what `value` would be on that path
v0 = loss.value(loss.to_device(P.x0))
eng.krylov_run(8, 10, False, v0, 1e-3, 0.5, 0.1)
loss.reset()
v0 = loss.value(loss.to_device(P.x0))
recs, _ = eng.krylov_run(5, 10, False, v0, 1e-3, 0.5, 0.1)
state = (recs[-1][0], recs[-1][12], recs[-1][2])
for tag, flush in (("kcrn_flush_l2", f_kcrn), ("synchronize only", f_none), ("kcrn_flush_l2 + sleep 2 ms", f_kcrn_sleep2)):
    ts, hs = [], []
    st = state
    for _ in range(20):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); t0 = time.perf_counter()
        recs, _ = eng.krylov_run(1, 10, False, st[0], st[1], 0.5, st[2])
        t1 = time.perf_counter(); b.record(); torch.cuda.synchronize()
        st = (recs[-1][0], recs[-1][12], recs[-1][2])
        ts.append(a.elapsed_time(b)); hs.append(1e3 * (t1 - t0))
    print(f"krylov_run(1) per launch behind {tag:30s} gpu {sum(ts) / len(ts):.3f} host {sum(hs) / len(hs):.3f} ms", flush=True)
