"""CPU: bv_sparse_encode16 (BV_CELLS_U32 -> BV_CELLS_U16 on the host).  Its vector loop (csrc/bv_encode16.cpp, AVX2, eight cells
per step) must write exactly the words of the scalar loop, and the words must decode back to the cells (the layout of
include/basevar_b200.h: gap | base << 5 | strand << 8 | phred << 9, "skip 31" words)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tile(seed, S, N, cov, with_odd=False):
    """Random sparse tile: (cells u32, aux u32, site_start).  Runs of neighbours, long gaps (skip words), sites of 0..N cells."""
    rng = np.random.default_rng(seed)
    cells, aux, start = [], [], [0]
    for s in range(S):
        k = rng.integers(0, 4)
        if k == 0:
            samp = np.nonzero(rng.random(N) < cov)[0]
        elif k == 1:
            samp = np.nonzero(rng.random(N) < cov * 0.05)[0]          # long gaps: several skip words in a row
        elif k == 2:
            a = int(rng.integers(0, N)); samp = np.arange(a, min(N, a + int(rng.integers(0, 40))))   # neighbours, gap 0
        else:
            samp = np.zeros(0, np.int64)
        base = rng.integers(0, 8, len(samp)); base[base == 5] = 0
        strand = rng.integers(0, 2, len(samp))
        if with_odd and len(samp) > 3 and s % 7 == 3:
            strand[int(rng.integers(0, len(samp)))] = 2                # no 16-bit form: the encoder must refuse
        phred = rng.integers(0, 94, len(samp))
        cells.append((samp | (base << 20) | (strand << 23) | (phred << 25)).astype(np.uint32))
        aux.append(rng.integers(0, 2 ** 24, len(samp)).astype(np.uint32))
        start.append(start[-1] + len(samp))
    return np.concatenate(cells), np.concatenate(aux), np.asarray(start, np.uint32)


def _encode(lib, cells, aux, start, slack):
    S = len(start) - 1
    cap = int(lib.bv_sparse_encode16_bound(len(cells), S, 1 << 20)) + slack
    w = np.full(cap + 8, 0xABCD, np.uint16); a16 = np.full(cap + 8, 0xDEADBEEF, np.uint32); so = np.zeros(S + 1, np.uint32)
    n = C.c_uint64(0)
    rc = lib.bv_sparse_encode16(cells.ctypes.data if len(cells) else None, aux.ctypes.data if aux is not None else None, start.ctypes.data, S,
                                w.ctypes.data, a16.ctypes.data if aux is not None else None, cap, so.ctypes.data, C.byref(n))
    assert (w[cap:] == 0xABCD).all() and (a16[cap:] == 0xDEADBEEF).all(), "wrote behind max_words"
    return rc, w[:n.value].copy(), a16[:n.value].copy(), so


def _decode(words, start16, S):
    out, st = [], [0]
    for s in range(S):
        nxt = 0
        for wd in words[start16[s]:start16[s + 1]]:
            wd = int(wd); gap = wd & 31
            if gap == 31:
                nxt += 31
                continue
            samp = nxt + gap
            out.append(samp | (((wd >> 5) & 7) << 20) | (((wd >> 8) & 1) << 23) | ((wd >> 9) << 25))
            nxt = samp + 1
        st.append(len(out))
    return np.asarray(out, np.uint32), np.asarray(st, np.uint32)


_CHILD = r"""
import sys, numpy as np
sys.path.insert(0, %r)
from basevar_b200 import capi
from tests.test_encode16_cpu import _tile, _encode
lib = capi.load_library()
res = {}
for seed, S, N, cov in ((1, 300, 1000, 0.1), (2, 200, 5000, 0.3), (3, 60, 100000, 0.1), (4, 400, 64, 0.9)):
    cells, aux, start = _tile(seed, S, N, cov)
    for with_aux in (False, True):
        rc, w, a16, so = _encode(lib, cells, aux if with_aux else None, start, 64)
        assert rc == 0
        res[f"w{seed}{int(with_aux)}"] = w; res[f"a{seed}{int(with_aux)}"] = a16; res[f"s{seed}{int(with_aux)}"] = so
np.savez(sys.argv[1], **res)
"""


def test_vector_loop_writes_the_scalar_loops_words(tmp_path):
    from basevar_b200 import build
    build.build()
    outs = []
    for tag, env in (("vec", {}), ("scalar", {"BASEVAR_B200_NO_AVX2": "1"})):
        path = str(tmp_path / f"{tag}.npz")
        subprocess.run([sys.executable, "-c", _CHILD % ROOT, path], check=True, env={**os.environ, **env}, cwd=ROOT)
        outs.append(np.load(path))
    assert sorted(outs[0].files) == sorted(outs[1].files)
    for k in outs[0].files:
        assert np.array_equal(outs[0][k], outs[1][k]), k


@pytest.mark.parametrize("slack", [0, 4, 64])   # 0 / 4: not enough room behind the worst case for 8-word stores -> the scalar loop runs
def test_words_decode_back_to_the_cells(slack):
    from basevar_b200 import build, capi
    build.build()
    lib = capi.load_library()
    for seed, S, N, cov in ((11, 250, 2000, 0.1), (12, 120, 30000, 0.2), (13, 300, 17, 0.7)):
        cells, aux, start = _tile(seed, S, N, cov)
        rc, w, a16, so = _encode(lib, cells, aux, start, slack)
        assert rc == capi.BV_OK
        back, st = _decode(w, so, S)
        assert np.array_equal(back, cells) and np.array_equal(st, start)
        # aux words travel with their cell, skip words carry 0
        is_skip = (w & 31) == 31
        assert np.array_equal(a16[~is_skip], aux) and not a16[is_skip].any()


def test_cells_without_a_16_bit_form_are_refused():
    from basevar_b200 import build, capi
    build.build()
    lib = capi.load_library()
    cells, aux, start = _tile(21, 200, 3000, 0.2, with_odd=True)
    rc, *_ = _encode(lib, cells, None, start, 64)
    assert rc == -1 and b"strand" in lib.bv_last_error(None)
    # descending samples inside a long site
    cells, aux, start = _tile(22, 50, 3000, 0.2)
    big = int(np.argmax(np.diff(start)))
    c0 = int(start[big])
    cells[c0 + 20], cells[c0 + 21] = cells[c0 + 21], cells[c0 + 20]
    rc, *_ = _encode(lib, cells, None, start, 64)
    assert rc == -1 and b"ascend" in lib.bv_last_error(None)
