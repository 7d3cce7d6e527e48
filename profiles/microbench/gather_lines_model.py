"""CPU model behind DESIGN.md section 3.1: distinct 128-byte lines per warp-wide gather (= L1TEX passes per load) of
the two sweeps, for the entries as stored, for gather-ordered tiles (what the library does; the real tile plan comes
from kcrn_plan_tiles, host only) and for a popularity renumbering of the gathered vector (next step, not built).

    python profiles/microbench/gather_lines_model.py [rcv1 w8a ...]     (no GPU needed)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
import conftest  # noqa: F401  (import paths)
from test_host_cpu import _plan_tiles
from kcrn_b200.synth import make_problem


def lines_per_warp(gather_idx, tiles, order_tiles):
    """mean distinct lines (16 doubles) over the warp loads of a sweep; thread t of a tile loads entries t, t+256, ..."""
    tot, loads = 0, 0
    for seg0, k0, z, w in tiles:
        nk, tag = w & 0xFFFF, w >> 16
        ix = gather_idx[k0:k0 + nk]
        if order_tiles and z > 0 and tag != 15:
            ix = np.sort(ix)
        pad = (-nk) % 32
        lines = np.concatenate((ix >> 4, np.full(pad, -1)))
        lines = np.sort(lines.reshape(-1, 32), axis=1)
        distinct = 1 + (np.diff(lines, axis=1) != 0).sum(axis=1) - (pad > 0) * (np.arange(lines.shape[0]) == lines.shape[0] - 1)
        tot += int(distinct.sum())
        loads += lines.shape[0]
    return tot / loads, tot


for name in sys.argv[1:] or ["rcv1", "w8a"]:
    P = make_problem(name, seed=0)
    A = P.scipy_csr()
    C = A.tocsc()
    # length-class order of the columns (the order the CSC copy is stored in)
    cnt = np.diff(C.indptr)
    cls = np.digitize(cnt, [9, 33, 129, 513, 2049])
    order = np.argsort(cls, kind="stable")
    pos = np.empty(P.d, np.int64)
    pos[order] = np.arange(P.d)
    csc_ptr = np.concatenate(([0], np.cumsum(cnt[order]))).astype(np.int32)
    csc_rows = np.concatenate([C.indices[C.indptr[c]:C.indptr[c + 1]] for c in order]).astype(np.int64)
    t_csr, _, it_csr = _plan_tiles(P.indptr)
    t_csc, _, it_csc = _plan_tiles(csc_ptr)
    cols = P.indices.astype(np.int64)
    print(f"{name}: n={P.n} d={P.d} nnz={P.nnz}; tiles CSR {len(t_csr)} x {it_csr * 256}, CSC {len(t_csc)} x {it_csc * 256}")
    for label, idx, tiles in (("CSR pass (gathers v[col])", cols, t_csr.tolist()), ("CSC pass (gathers t[row])", csc_rows, t_csc.tolist())):
        a, ta = lines_per_warp(idx, tiles, False)
        b, tb = lines_per_warp(idx, tiles, True)
        print(f"  {label}: stored order {a:5.2f} lines per warp load ({ta} passes), gather-ordered tiles {b:5.2f} ({tb})")
    c, tc = lines_per_warp(pos[cols], t_csr.tolist(), True)
    print(f"  CSR pass with the vector kept in the CSC copy's column order (not built): {c:5.2f} ({tc})")
