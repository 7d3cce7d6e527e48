"""Sweep and Lanczos-graph timing of one libkcrn variant (KCRN_LIBRARY), rcv1 / w8a shapes."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
import conftest  # noqa
import torch
from kcrn_b200.engine import Engine
from kcrn_b200.synth import make_problem
from oracle import oracle as O
for name in sys.argv[1:] or ["rcv1"]:
    P = make_problem(name, seed=0)
    eng = Engine(P.indptr, P.indices, P.data, O.map_labels(P.b), (P.n, P.d))
    eng.set_x(torch.as_tensor(P.x0, device="cuda"))
    eng.gradient(keep_internal=True)
    us = [min(eng.time_sweep(w, 200) for _ in range(3)) for w in (0, 1)]
    try:
        alt = min(eng.time_sweep(2, 200) for _ in range(3))
        print(f"   plain passes alternating (rows, cols, rows, ...): {alt:6.2f} us per pass, {2 * alt:6.2f} per plain HVP; each pass repeated on its own: {us[0] + us[1]:6.2f}")
    except Exception as e:
        print("   alternating passes: not in this build", str(e)[:60])
    for _ in range(5):
        eng.lanczos(10, False)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    allb = []
    for _ in range(5):
        a.record()
        for _ in range(50):
            eng.lanczos(10, False)
        b.record(); torch.cuda.synchronize()
        allb.append(round(a.elapsed_time(b) / 50 * 1e2, 2))
        best = min(best, a.elapsed_time(b) / 50 * 1e3)
    tag = os.path.basename(os.environ.get('KCRN_LIBRARY', 'default')) + " KCRN_GATHER=" + os.environ.get("KCRN_GATHER", "auto")
    if hasattr(eng, "gather_variants") and hasattr(eng.lib, "kcrn_gather_variants"):
        tag += " -> " + ",".join(str(v) for v in eng.gather_variants().values())
    print(f"{tag:46s} {name:6s} sweeps {us[0]:6.2f} {us[1]:6.2f} us   Lanczos graph {best / 10:6.2f} us per HVP  batches {allb}")
    x = eng.new_vector(); x.copy_(torch.as_tensor(P.x0, device="cuda")); x.mul_(0.2)
    eng.set_x(x); eng.gradient(keep_internal=True)
    a.record()
    for _ in range(20):
        eng.lanczos(10, False)
    b.record(); torch.cuda.synchronize()
    print("   at another x, 20 reps once:", round(a.elapsed_time(b) / 20 * 1e2, 2), "us per HVP, hvps", int(eng.lanczos_view()[3][3].item()))
    eng.close()
