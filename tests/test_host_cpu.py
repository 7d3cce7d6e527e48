"""CPU: the C-ABI library loads and exports every symbol include/kcrn.h declares (no compute calls),
the ctypes table matches the header, and the host-side mirror (labels, run loop, trace) behaves like
the reference's host code."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "kcrn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kcrn_[a-z0-9_A-Z]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from kcrn_b200 import _lib, build
    so = build.build()
    assert os.path.exists(so)
    lib = ctypes.CDLL(so)
    names = _header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/kcrn.h but not exported by libkcrn.so"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    assert _lib.load().kcrn_abi_version() == 1


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    from kcrn_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(np.zeros(2, np.int32), np.zeros(0, np.int32), np.zeros(0), np.zeros(1), (1, 4))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "krylov-cubic-regularized-newton_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(base, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "kcrn_oracle" not in src, f


def test_label_mapping_matches_reference(golden):
    from optimizer.loss import map_labels
    g = golden("prims")
    for i in range(5):
        assert np.array_equal(map_labels(g[f"label_in{i}"]), g[f"label_out{i}"])
    with pytest.raises(ValueError):
        map_labels(np.array([0., 1., 2.]))
    # row shards take the encoding of the GLOBAL labels: a single-class shard and a shard starting with the other label
    from optimizer.loss import label_encoding
    b = np.array([3., 5., 5., 3., 5., 5.])
    enc = label_encoding(b)
    assert np.array_equal(np.concatenate([map_labels(b[:4], enc), map_labels(b[4:], enc)]), map_labels(b))
    b = np.array([-1., 1., 1., 1.])
    assert np.array_equal(map_labels(b[1:], label_encoding(b)), map_labels(b)[1:])


class _QuadLoss:
    """Stand-in oracle for the run loop: f(x) = |x|^2 / 2."""
    regularizer = None
    f_opt, x_opt, n = np.inf, None, 1

    def value(self, x):
        v = 0.5 * float(np.dot(x, x))
        self.f_opt = min(self.f_opt, v)
        return v

    @staticmethod
    def norm(x):
        return float(np.linalg.norm(x))

    def reset(self):
        pass


def test_run_loop_stopping_rules_and_trace():
    from optimizer.optimizer import Optimizer

    class Halve(Optimizer):
        def step(self):
            self.x = self.x * 0.5

    loss = _QuadLoss()
    opt = Halve(loss=loss, tolerance=1e-3, label="h", tqdm=False)
    tr = opt.run(np.ones(4), it_max=1000)
    # |x_k - x_{k-1}| = 2 * 0.5^k: first below 1e-3 after 11 steps; the loop notices before step 12
    assert opt.it == 11 and tr.its[:6] == [0, 1, 2, 3, 4, 5] and len(tr.xs) == len(tr.ts) == len(tr.its)
    opt.compute_loss_of_iterates()
    assert np.all(np.diff(tr.loss_vals) < 0) and tr.best_loss_value == tr.loss_vals[-1]
    opt2 = Halve(loss=loss, tolerance=0, label="h", tqdm=False)
    opt2.run(np.ones(4), it_max=7)
    assert opt2.it == 7
    opt3 = Halve(loss=loss, label="h", tqdm=False)
    opt3.run(np.ones(4), t_max=0.05)
    assert opt3.it > 0 and opt3.trace.ts[-1] <= 0.2


def test_trace_pickle_roundtrip(tmp_path):
    from optimizer.opt_trace import Trace
    loss = _QuadLoss()
    tr = Trace(loss, label="t")
    tr.xs, tr.ts, tr.its = [np.ones(3), np.zeros(3)], [0, 0.1], [0, 1]
    tr.compute_loss_of_iterates()
    tr.save("t.pkl", path=str(tmp_path) + "/")
    back = Trace.from_pickle(os.path.join(str(tmp_path), "t.pkl"), loss=loss)
    assert back.loss is loss and np.array_equal(back.loss_vals, tr.loss_vals) and tr.loss is loss


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` is the CPU arm the driver times beside the GPU arm: it must run without a GPU and
    print ONE JSON line with the contract's keys (metric / unit / config of the GPU arm, cpu_baseline, e2e)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "small",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "krylov_crn_hvps_per_sec" and line["unit"] == "HVP/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] == 2
    installed = os.path.isdir(os.path.join(root, "baseline", "_ref", "optimizer"))
    assert line["cpu_baseline"]["kind"] == ("reference" if installed else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]


def test_bench_roofline_traffic_comes_from_the_committed_capture():
    import json
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    with open(os.path.join(root, "profiles", "ncu_traffic.json")) as f:
        t = json.load(f)
    assert bench.ncu_traffic(t["workload"], 0) == t["csr_pass_bytes_per_launch"]
    assert bench.ncu_traffic(t["workload"], 1) == t["csc_pass_bytes_per_launch"]
    assert bench.ncu_traffic("no-such-workload", 0) is None
    assert os.path.exists(os.path.join(root, t["source"].split(" ")[0]))


def _plan_tiles(indptr):
    from ctypes import byref, c_int, c_int64
    from kcrn_b200 import _lib
    lib = _lib.load()
    indptr = np.ascontiguousarray(indptr, np.int32)
    nseg = indptr.size - 1
    nt, nl, items = c_int64(), c_int64(), c_int()
    _lib.check(lib.kcrn_plan_tiles(indptr.ctypes.data, nseg, None, 0, byref(nt), None, 0, byref(nl), byref(items)))
    tiles = np.zeros((max(nt.value, 1), 4), np.int32)
    longtab = np.zeros((max(nl.value, 1), 4), np.int32)
    _lib.check(lib.kcrn_plan_tiles(indptr.ctypes.data, nseg, tiles.ctypes.data, tiles.shape[0], byref(nt),
                                   longtab.ctypes.data, longtab.shape[0], byref(nl), byref(items)))
    return tiles[:nt.value], longtab[:nl.value], items.value


def _check_tiling(lengths):
    lengths = np.asarray(lengths, np.int64)
    indptr = np.concatenate(([0], np.cumsum(lengths))).astype(np.int32)
    nseg, nnz = lengths.size, int(indptr[-1])
    tiles, longtab, items = _plan_tiles(indptr)
    assert items in (8, 11, 12)
    tile_nnz = items * 256
    assert tile_nnz <= 4096                                   # natural positions fit the 12 slot bits of a gather word
    seg_cover = np.zeros(nseg, np.int64)
    nnz_cover = np.zeros(nnz, np.int64)
    pieces_seen = {}
    for seg0, k0, z, w in tiles.tolist():
        nk, tag = w & 0xFFFF, w >> 16
        assert 0 <= nk <= tile_nnz and 0 <= k0 and k0 + nk <= nnz
        nnz_cover[k0:k0 + nk] += 1
        if z > 0 and tag != 15:                               # stream tile: z whole segments
            assert z <= 1024 and 0 <= tag <= 5
            assert k0 == indptr[seg0] and nk == indptr[seg0 + z] - indptr[seg0]
            seg_cover[seg0:seg0 + z] += 1
        elif z > 0:                                           # one medium segment
            assert z == 1 and k0 == indptr[seg0] and nk == lengths[seg0]
            seg_cover[seg0] += 1
        else:                                                 # one piece of a long segment
            slot0, npieces, first, seg = longtab[-z - 1].tolist()
            assert seg == seg0 and first == indptr[seg0] and lengths[seg0] > tile_nnz
            assert npieces == -(-int(lengths[seg0]) // tile_nnz)
            assert (k0 - first) % tile_nnz == 0 and nk == min(tile_nnz, indptr[seg0 + 1] - k0)
            pieces_seen.setdefault(seg0, []).append((k0 - first) // tile_nnz)
    for seg, got in pieces_seen.items():
        slot0, npieces, first, s = longtab[[i for i in range(len(longtab)) if longtab[i, 3] == seg][0]].tolist()
        assert sorted(got) == list(range(npieces))
        seg_cover[seg] += 1
    assert np.all(nnz_cover == 1) and np.all(seg_cover == 1)  # every entry and every segment (empty ones too) exactly once
    slots = sorted((int(r[0]), int(r[1])) for r in longtab)   # partial-sum slots of the long segments do not overlap
    for (a, na), (b, _) in zip(slots, slots[1:]):
        assert a + na <= b


def test_tile_plan_invariants_on_awkward_length_distributions():
    rng = np.random.default_rng(0)
    _check_tiling([0])
    _check_tiling([5])
    _check_tiling(np.zeros(3000, np.int64))                               # only empty segments
    _check_tiling([4097, 0, 0, 1, 20000, 3, 2816, 2817, 3072, 3073, 0])   # around every threshold
    _check_tiling(rng.poisson(74, 20242))                                  # rcv1-like rows
    _check_tiling(np.clip((rng.pareto(0.9, 47236) * 3).astype(np.int64), 0, 20000))   # power-law columns
    _check_tiling(rng.integers(0, 3, 200_000))                             # many tiny segments (> 1024 per tile's worth)
    _check_tiling(np.concatenate((rng.poisson(24, 49749), [49749] * 3)))   # w8a-like with three dense columns


def test_tile_plan_invariants_hypothesis():
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")

    @hyp.settings(max_examples=60, deadline=None)
    @hyp.given(st.lists(st.one_of(st.integers(0, 40), st.integers(0, 1200), st.integers(2000, 9000)), min_size=1,
                        max_size=400))
    def run(lengths):
        _check_tiling(lengths)

    run()


# ------------------------------------------------------------------ balanced one-tile-per-SM plan (kcrn_persist.cu)
def _plan_persist(indptr, grid):
    from ctypes import byref, c_int, c_int64
    from kcrn_b200 import _lib
    lib = _lib.load()
    indptr = np.ascontiguousarray(indptr, np.int32)
    nseg = indptr.size - 1
    nt, nl, fits = c_int64(), c_int64(), c_int()
    _lib.check(lib.kcrn_plan_persist(indptr.ctypes.data, nseg, grid, None, 0, byref(nt), None, 0, byref(nl), byref(fits)))
    tiles = np.zeros((max(nt.value, 1), 12), np.int32)
    longtab = np.zeros((max(nl.value, 1), 4), np.int32)
    _lib.check(lib.kcrn_plan_persist(indptr.ctypes.data, nseg, grid, tiles.ctypes.data, tiles.shape[0], byref(nt),
                                     longtab.ctypes.data, longtab.shape[0], byref(nl), byref(fits)))
    return tiles[:nt.value], longtab[:nl.value], bool(fits.value)


def _check_persist_plan(lengths, grid=148, must_fit=None):
    lengths = np.asarray(lengths, np.int64)
    indptr = np.concatenate(([0], np.cumsum(lengths))).astype(np.int32)
    nseg, nnz = lengths.size, int(indptr[-1])
    tiles, longtab, fits = _plan_persist(indptr, grid)
    if must_fit is not None:
        assert fits == must_fit
    if not fits:
        return None
    assert 1 <= len(tiles) <= grid
    seg_whole = np.zeros(nseg, np.int64)
    pieces = {}
    e = 0
    for (e0, nk, seg0, ns, lg, hl, hp, tl, tp, *_pad) in tiles.tolist():
        assert e0 == e and 0 <= nk <= 12288 and 1 <= ns <= 4096 and 0 <= lg <= 5
        e += nk
        local = 0
        for i in range(ns):
            sg = seg0 + i
            a, b = max(int(indptr[sg]), e0), min(int(indptr[sg + 1]), e0 + nk)
            ln = max(0, b - a)
            local += ln
            cut = (i == 0 and hl >= 0) or (i == ns - 1 and tl >= 0)
            if cut:
                l, p = (hl, hp) if (i == 0 and hl >= 0) else (tl, tp)
                if i == 0 and hl >= 0 and i == ns - 1 and tl >= 0:
                    assert (hl, hp) == (tl, tp)                 # a middle piece: the whole tile inside one segment
                assert longtab[l][2] == sg and lengths[sg] > 1024 and ln > 0
                pieces.setdefault(l, []).append((p, a, b))
            else:
                assert a == indptr[sg] and b == indptr[sg + 1], "an uncut segment lies whole inside its tile"
                seg_whole[sg] += 1
        assert local == nk
    assert e == nnz
    slots = []
    for l, (slot0, npieces, sg, _z) in enumerate(longtab.tolist()):
        got = sorted(pieces[l])
        assert [p for p, _, _ in got] == list(range(npieces)) and npieces >= 2
        assert got[0][1] == indptr[sg] and got[-1][2] == indptr[sg + 1]
        for (_, _, b0), (_, a1, _) in zip(got, got[1:]):
            assert b0 == a1                                     # the pieces tile the segment in order
        assert seg_whole[sg] == 0
        seg_whole[sg] += 1
        slots.append((slot0, npieces))
    assert np.all(seg_whole == 1)                               # every segment (empty ones too) exactly once
    slots.sort()
    for (a, na), (b, _) in zip(slots, slots[1:]):
        assert a + na <= b
    sizes = tiles[:, 1]
    return int(sizes.max()), int(sizes.min())


def test_persist_plan_invariants():
    rng = np.random.default_rng(1)
    assert _check_persist_plan([5]) is not None
    assert _check_persist_plan([0, 0, 3, 0]) is not None
    _check_persist_plan(np.zeros(10, np.int64), must_fit=False)                         # no entries: per-pass kernels
    mx, mn = _check_persist_plan(rng.poisson(74, 20242), must_fit=True)                 # rcv1-like rows
    assert mx - mn <= 1024 and mx <= 10132 + 600
    cols = np.sort(np.clip((1.33e5 / np.arange(1, 47237)).astype(np.int64), 0, None))   # Zipf columns, shortest first
    mx, mn = _check_persist_plan(cols, must_fit=True)
    _check_persist_plan(np.concatenate((rng.poisson(24, 49749), [49749] * 3)), must_fit=True)   # w8a-like + dense columns
    _check_persist_plan([4000] * 300, must_fit=True)                                    # w8a-like columns: all cut
    _check_persist_plan(rng.integers(0, 3, 400_000), must_fit=True)                     # tiny segments: the 4096-segment cap
    _check_persist_plan(rng.integers(0, 3, 1_400_000), must_fit=False)                  # ... which no longer fits 148 tiles
    _check_persist_plan(rng.poisson(74, 40000), must_fit=False)                         # 3 M entries: too large
    _check_persist_plan([1_600_000], must_fit=True)                                     # one segment cut into 148 pieces
    _check_persist_plan([1025, 1024, 1023] * 500, grid=7, must_fit=False)
    for g in (1, 2, 5, 148):
        _check_persist_plan(rng.integers(0, 2000, 40), grid=g)


def test_persist_plan_hypothesis():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.one_of(st.integers(0, 40), st.integers(0, 3000), st.just(0), st.integers(1000, 30000)),
                    min_size=1, max_size=300), st.integers(1, 40))
    def run(lengths, grid):
        _check_persist_plan(lengths, grid=grid)

    run()


def test_best_point_bookkeeping_of_the_oracle():
    """Oracle.value's best-point tracking (loss.py:70-72): x_opt is a copy taken when the value was noted.  For a trial
    point the optimizer owns the copy is deferred to the first read of x_opt -- same content, a distinct tensor, made once;
    an in-place write to the noted tensor before that read fails loudly instead of handing out the wrong point; every
    other writer (plain assignment, numpy iterates, a better value found later) replaces a deferred copy."""
    import torch
    from optimizer.loss import Oracle
    o = Oracle()
    assert o.x_opt is None and o.f_opt == np.inf
    x = torch.arange(4, dtype=torch.float64)
    o._note_value(1.0, x, owned=True)
    got = o.x_opt
    assert o.f_opt == 1.0 and torch.equal(got, x) and got is not x and o.x_opt is got
    o._note_value(2.0, torch.zeros(4, dtype=torch.float64), owned=True)        # not better: nothing changes
    assert o.f_opt == 1.0 and o.x_opt is got
    y = torch.ones(4, dtype=torch.float64)
    o._note_value(0.5, y, owned=True)
    y.add_(1.0)
    with pytest.raises(RuntimeError, match="modified in place"):
        o.x_opt
    z = np.array([3.0, 4.0])
    o._note_value(0.25, z)                                                     # numpy iterate: deep copy right away
    z[0] = -1.0
    assert np.array_equal(o.x_opt, [3.0, 4.0]) and o.f_opt == 0.25
    w = torch.full((4,), 7.0, dtype=torch.float64)
    o._note_value(0.125, w)                                                    # not owned: cloned right away
    w.zero_()
    assert torch.equal(o.x_opt, torch.full((4,), 7.0, dtype=torch.float64))
    o.x_opt = None
    assert o.x_opt is None


class _OracleEngine:
    """Stand-in for kcrn_b200.engine.Engine on a machine without a GPU: the two calls Cubic_Krylov_LS.step makes, answered
    by the CPU oracle (test infrastructure), so that the HOST side of the step -- which M is tried first, the acceptance
    test, M / beta on rejection, the commit, the counters, the best point -- can be checked against the oracle's own loop."""
    world, p2p = 1, False

    def __init__(self, L, d):
        self.L, self.d, self.x, self.best = L, d, None, None
        self.trials = 0

    def new_vector(self, size=None):
        return np.empty(self.d if size is None else size)

    def _first(self, m, reorth):
        from oracle import oracle as O
        x = self.x
        g = self.L.gradient(x)
        self.V, al, be, _ = O.lanczos(lambda v: self.L.hess_vec_prod(x, v), g, m, reorth)
        self.T = np.diag(al) + np.diag(be, -1) + np.diag(be, 1)
        self.gs = np.zeros(al.size)
        self.gs[0] = self.gnorm = np.linalg.norm(g)
        self.hvps = m

    def krylov_trial(self, first, m, reorth, M, r0, eps=1e-8, it_max=100, x_out=None):
        from oracle import oracle as O
        if first:
            self._first(m, reorth)
        self.trials += 1
        s, its, r, dec = O.cubic_solver_root(self.gs, self.T, M, it_max=it_max, epsilon=eps, r0=r0)
        xn = self.x + self.V @ s
        x_out[:] = xn
        self.x_trial = xn
        return self.L.value(xn), dec, r, its, 0, self.gs.size, self.gnorm, self.hvps, float(np.linalg.norm(xn - self.x))

    def krylov_step(self, m, reorth, value0, reg_coef, beta, r0, eps, solver_it_max, max_tries, f_opt):
        """What kcrn_krylov_run does for one iteration: the line search of cubic.py:286-302 and the 16-double record."""
        M, tries, first, x0 = reg_coef * beta, 0, True, self.x
        while True:
            xn = self.new_vector()
            f, dec, r, its, _, m_eff, gnorm, hv, dx = self.krylov_trial(first, m, reorth, M, r0, eps, solver_it_max, x_out=xn)
            first = False
            if f < f_opt:
                f_opt, self.best = f, xn.copy()
            if not (f > value0 - dec and tries < max_tries):
                break
            M, tries = M / beta, tries + 1
        self.x = xn
        rec = [f, dec, r, its, 0, m_eff, gnorm, hv, dx, 0, 0, f_opt, M, tries, 0.0, 1.0]
        return rec, 0, xn.copy()

    def best_point(self):
        return self.best.copy()


def _stub_loss(P):
    from optimizer.loss import Oracle
    from oracle import oracle as O

    class StubLoss(Oracle):
        def __init__(self):
            super().__init__(l1=0, l2=0)
            self.L = O.OracleLogistic(P.indptr, P.indices, P.data, (P.n, P.d), P.b)
            self.engine = _OracleEngine(self.L, P.d)
            self.n, self.dim = P.n, P.d
            self.calls = {"value": 0, "gradient": 0, "hess_vec_prod": 0, "forward": 0}

        def to_device(self, x):
            return np.array(x, dtype=np.float64)

        def _value(self, x):
            self.engine.x = np.array(x)
            return self.L.value(x)

        def _forward(self, x):
            self.engine.x = np.array(x)

        def adopt_iterate(self, x):
            pass

        def forget_iterate(self):
            pass

        def reset(self):
            pass

        @staticmethod
        def norm(x):
            return float(np.linalg.norm(x))
    return StubLoss()


@pytest.mark.parametrize("reg_coef", [1e-7, 1e-11])     # 12 rejections in the first step / the cap of 20 rejections
@pytest.mark.parametrize("device_step", [False, True])
def test_krylov_step_host_logic_against_the_oracle_loop(device_step, reg_coef):
    """Cubic_Krylov_LS.step with the engine replaced by the oracle: every accepted value, M, backtrack count, Newton
    iteration count and the best point must be the oracle loop's (oracle.KrylovCRN, cubic.py:265-309) exactly -- the same
    arithmetic runs on both sides, only the control flow under test differs.  Both step paths of the mirror."""
    from kcrn_b200.synth import make_problem
    from optimizer.cubic import Cubic_Krylov_LS
    from oracle import oracle as O
    P = make_problem("tiny", seed=3)
    loss = _stub_loss(P)
    opt = Cubic_Krylov_LS(loss=loss, reg_coef=reg_coef, subspace_dim=6, tolerance=0, label="s", tqdm=False,
                          device_step=device_step)
    opt.run(P.x0, it_max=12)
    ora = O.KrylovCRN(O.OracleLogistic(P.indptr, P.indices, P.data, (P.n, P.d), P.b), reg_coef=reg_coef, subspace_dim=6,
                      tolerance=0)
    ora.run(P.x0, it_max=12)
    assert len(opt.history) == len(ora.history) == 12
    for a, b in zip(opt.history, ora.history):
        assert (a["f"], a["M"], a["backtracks"], a["solver_it"], a["m_eff"]) == \
               (b["f"], b["M"], b["backtracks"], b["solver_it"], b["m_eff"]), (a, b)
    assert ora.history[0]["backtracks"] == (12 if reg_coef == 1e-7 else 20)      # the line search rejected / hit its cap
    assert np.array_equal(opt.x, ora.x) and opt.reg_coef == ora.M and opt.r0 == ora.r0
    trials = 12 + sum(h["backtracks"] for h in ora.history)
    assert loss.engine.trials == trials
    assert loss.calls["gradient"] == 12 and loss.calls["hess_vec_prod"] == 12 * 6
    assert loss.calls["value"] == trials == loss.calls["forward"]        # one value and one forward sweep per trial
    assert loss.f_opt <= min(h["f"] for h in ora.history)                # rejected trial points count as well (loss.py:70-72)
    assert loss.x_opt is not None and abs(loss.L.value(np.asarray(loss.x_opt)) - loss.f_opt) <= 1e-15
