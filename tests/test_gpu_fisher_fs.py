"""-m gpu: bv_fisher_fs -- FS (-10 log10 of the two-sided Fisher exact p, src/basetype.cpp:277-283, htslib/kfunc.c:245-313) of
free-standing 2x2 strand tables on the device, against the compiled reference's values (tests/golden/fisher_fs.npz, made by
tests/golden/make_golden.py from strand_bias()) and against the oracle on random tables: narrow supports (the reference's own
walk), wide ones (bisection + tail sums), margins of 1, empty rows."""
import ctypes as C
import math
import os

import numpy as np
import pytest

import basevar_b200 as bv
from basevar_b200 import capi
from tests import util

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _fs(eng, tables):
    t = np.ascontiguousarray(tables, np.int32)
    out = np.zeros(len(t), np.float64)
    eng._check(eng.lib.bv_fisher_fs(eng._ctx, t.ctypes.data, len(t), out.ctypes.data), "bv_fisher_fs")
    return out


def _fs_of_p(p):   # src/basetype.cpp:277-283
    fs = -10 * math.log10(p) if p > 0 else float("inf")
    if math.isinf(fs):
        return 10000.0
    return 0.0 if fs == 0 else fs


def test_reference_fixture(built_lib):
    d = np.load(os.path.join(HERE, "golden", "fisher_fs.npz"))
    eng = bv.BaseTypeEngine(device=0, max_samples=int(d["tables"].sum(axis=1).max()), min_af=0.01)
    try:
        got = _fs(eng, d["tables"])
    finally:
        eng.close()
    assert util.close(got, d["fs"], util.RTOL, util.FS_ATOL).all(), (got, d["fs"])


def test_random_tables_against_the_oracle(built_lib, oracle_lib):
    rng = np.random.default_rng(2024)
    tables = []
    for scale in (3, 12, 60, 400, 3000, 20000):
        t = rng.integers(0, scale, size=(300, 4))
        t[::7, 2:] = rng.integers(0, 3, size=(len(t[::7]), 2))      # a small ALT row: margins of 1 and 2
        t[::11, :2] = 0                                              # empty REF row
        t[::13, 2:] = 0                                              # empty ALT row
        tables.append(t)
    tables = np.concatenate(tables).astype(np.int32)
    eng = bv.BaseTypeEngine(device=0, max_samples=int(tables.sum(axis=1).max()), min_af=0.01)
    try:
        got = _fs(eng, tables)
        with pytest.raises(bv.BvError, match="max_samples"):
            _fs(eng, np.array([[1 << 20, 1 << 20, 5, 5]], np.int32))
        with pytest.raises(bv.BvError, match="negative"):
            _fs(eng, np.array([[1, -1, 5, 5]], np.int32))
    finally:
        eng.close()
    want = np.empty(len(tables))
    atol = np.full(len(tables), util.FS_ATOL)
    for i, (a, b, c, d) in enumerate(tables.tolist()):
        # strand_bias() runs the test on every table (an empty row gives p == 1)
        p = oracle_lib.bvo_fisher_two_sided(a, b, c, d)
        want[i] = _fs_of_p(p)
        if 0 < p < 2.3e-308:
            # A p-value in the denormal range is a sum of hypergeometric terms that are each a few hundred to a few thousand steps
            # of the denormal grid (4.9e-324): the reference's own sum is that coarse, and below ~4,000 steps its very first term
            # -- exp() of the observed table's log-probability -- is 0 or not by the last bit of exp (device: FS = 10000, the
            # reference's rule for p == 0).  Measured on a ladder of such tables: tools/fs_probe.py, profiles/r02_fs_probe.txt.
            steps = p / 4.94e-324
            atol[i] = 10 / math.log(10) * 4096 / steps if steps > 4096 else np.inf
    ok = (np.abs(got - want) <= atol + util.RTOL * np.abs(want)) | (np.isinf(atol) & ((got == 10000.0) | (np.abs(got - want) < 5)))
    bad = np.nonzero(~ok)[0][:10]
    assert ok.all(), (bad, tables[bad], got[bad], want[bad])
