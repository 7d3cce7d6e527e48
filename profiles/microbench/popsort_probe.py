"""What would the recurrence gain if the d-vectors lived in column-popularity order?  Same rcv1-shaped matrix, columns
relabelled before it reaches the engine (no engine change): natural (random popularity permutation), descending count,
ascending length class.  Sweeps + Lanczos graph per HVP."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
import conftest  # noqa
import numpy as np
import torch
from kcrn_b200.engine import Engine
from kcrn_b200.synth import make_problem
from oracle import oracle as O


def run(tag, A, b, x0):
    A.sort_indices()
    eng = Engine(A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data, O.map_labels(b), A.shape)
    eng.set_x(torch.as_tensor(x0, device="cuda"))
    eng.gradient(keep_internal=True)
    us = [min(eng.time_sweep(w, 200) for _ in range(3)) for w in (0, 1)]
    for _ in range(5):
        eng.lanczos(10, False)
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        a.record()
        for _ in range(50):
            eng.lanczos(10, False)
        e.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(e) / 50 * 1e3)
    gv = ",".join(str(v) for v in eng.gather_variants().values())
    print(f"{tag:28s} variants {gv}  sweeps {us[0]:6.2f} {us[1]:6.2f} us   Lanczos graph {best / 10:6.2f} us per HVP", flush=True)
    eng.close()


for name in sys.argv[1:] or ["rcv1"]:
    P = make_problem(name, seed=0)
    A = P.scipy_csr()
    cnt = np.bincount(A.indices, minlength=P.d)
    run(name + " natural", A.copy(), P.b, P.x0)
    order = np.argsort(-cnt, kind="stable")             # new position q holds old column order[q]
    run(name + " descending count", A[:, order].tocsr(), P.b, P.x0)
    cls = np.digitize(cnt, [9, 33, 129, 513])
    order = np.argsort(cls, kind="stable")
    run(name + " ascending class", A[:, order].tocsr(), P.b, P.x0)
    order = np.argsort(-cls, kind="stable")
    run(name + " descending class", A[:, order].tocsr(), P.b, P.x0)
