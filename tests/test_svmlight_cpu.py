"""The host-side LIBSVM / svmlight parser against sklearn.datasets.load_svmlight_file -- the call the reference
driver makes (cubic_newton.py:52) -- bit for bit: labels, indptr, indices, data, shape."""
import bz2
import gzip
import os

import numpy as np
import pytest

from kcrn_b200.synth import make_problem


def _same(X, y, Xs, ys):
    assert X.shape == Xs.shape and X.dtype == Xs.dtype == np.float64
    assert np.array_equal(y, ys)
    assert np.array_equal(X.indptr, Xs.indptr) and np.array_equal(X.indices, Xs.indices)
    assert np.array_equal(X.data.view(np.uint64), Xs.data.view(np.uint64))


@pytest.mark.parametrize("name,zero_based", [("tiny", True), ("tiny", False), ("small", False), ("w8a", False)])
def test_dumped_problem_round_trip_equals_sklearn(tmp_path, name, zero_based):
    from sklearn.datasets import dump_svmlight_file, load_svmlight_file as sk_load
    from kcrn_b200.svmlight import load_svmlight_file
    P = make_problem(name, seed=1)
    path = str(tmp_path / f"{name}.svm")
    dump_svmlight_file(P.scipy_csr(), P.b, path, zero_based=zero_based)
    Xs, ys = sk_load(path)
    for threads in (1, 5, 0):
        X, y = load_svmlight_file(path, threads=threads)
        _same(X, y, Xs, ys)
    assert np.allclose(X.data, P.scipy_csr().data, rtol=1e-15, atol=0)      # the dump keeps 16 significant digits


def test_grammar_corner_cases_equal_sklearn(tmp_path):
    from sklearn.datasets import load_svmlight_file as sk_load
    from kcrn_b200.svmlight import load_svmlight_file
    text = (b"# header comment\n"
            b"+1 1:2e-3 3:1.5 7:-0.25   # trailing comment 9:9\n"
            b"\n"
            b"-1 qid:4 2:1 5:0 6:1E+2\r\n"
            b"   \t\n"
            b"1.5\n"                                  # a row without features
            b"-1 4:.5 10:0.1 11:7.\n"
            b"0 1:1e-320 2:1.7976931348623157e308")   # denormal, largest double, no final newline
    path = str(tmp_path / "corner.svm")
    open(path, "wb").write(text)
    _same(*load_svmlight_file(path), *sk_load(path))
    for zb in (True, False):
        _same(*load_svmlight_file(path, zero_based=zb), *sk_load(path, zero_based=zb))
    _same(*load_svmlight_file(path, n_features=40), *sk_load(path, n_features=40))
    with pytest.raises(ValueError):
        load_svmlight_file(path, n_features=3)
    # compressed by extension, like sklearn
    for opener, ext in ((gzip.open, ".gz"), (bz2.open, ".bz2")):
        with opener(path + ext, "wb") as f:
            f.write(text)
        _same(*load_svmlight_file(path + ext), *sk_load(path + ext))


def test_zero_index_makes_auto_zero_based_and_cache_round_trip(tmp_path):
    from sklearn.datasets import load_svmlight_file as sk_load
    from kcrn_b200.svmlight import load_svmlight_file
    path = str(tmp_path / "zb.svm")
    open(path, "wb").write(b"1 0:1 3:2\n-1 1:4\n")
    _same(*load_svmlight_file(path), *sk_load(path))
    first = load_svmlight_file(path, cache=True)
    assert os.path.exists(path + ".kcrn.npz")
    _same(*load_svmlight_file(path, cache=True), *first)


def test_malformed_lines_fail_loudly_with_the_line_number(tmp_path):
    from kcrn_b200.svmlight import parse_svmlight
    for bad in (b"1 2:3\nx 1:1\n", b"1 2:3\n1 a:1\n", b"1 2:3\n1 4:\n1 1:1\n", b"1 2:3\n1 -4:1\n", b"1 2:3\n1 4 5\n"):
        with pytest.raises(ValueError, match="line 2"):
            parse_svmlight(bad)
    for bad in (b"1 2:3\n-1 10:0.1 4:.5\n", b"1 2:3\n-1 4:.5 4:7\n"):          # sklearn: "should be sorted and unique"
        with pytest.raises(ValueError, match="sorted and unique .line 2"):
            parse_svmlight(bad)
    labels, indptr, indices, data, lo, hi = parse_svmlight(b"")
    assert len(labels) == 0 and list(indptr) == [0] and len(data) == 0
