"""CPU (-m "not gpu"): the C-ABI library builds for sm_100a, loads, and exports every symbol include/basevar_b200.h
declares; entry points that need a device fail loudly (no CPU fallback); the host twin of the synthetic generator."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import basevar_b200 as bv
from basevar_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "basevar_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(bv_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree(built_lib):
    assert header_symbols() == sorted(capi.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(built_lib):
    lib = C.CDLL(built_lib)
    for name in header_symbols():
        assert hasattr(lib, name), f"{name} is declared in include/basevar_b200.h but not exported"
    assert capi.load_library().bv_version() == 1


def test_struct_layouts_match_the_header(built_lib):
    assert capi.SITE_OUT_DTYPE.itemsize == 128
    assert C.sizeof(capi.BvParams) == 36 and C.sizeof(capi.BvTile) == 56 and C.sizeof(capi.BvSparseTile) == 56
    assert C.sizeof(capi.BvSynthModel) == 32 + 4 * (96 + 1024 + 256)
    offs = {n: capi.SITE_OUT_DTYPE.fields[n][1] for n in capi.SITE_OUT_DTYPE.names}
    assert offs == {"depth": 0, "depth_other": 16, "reserved0": 20, "fwd": 24, "rev": 40, "n_alt": 56, "alt": 57,
                    "n_active": 61, "flags": 62, "em_calls": 63, "af": 64, "qual": 96, "chi2": 104, "fs_cvg": 112,
                    "fs_vcf": 120}


def test_no_cpu_fallback(built_lib):
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(bv.BvError, match="no CUDA device|CUDA"):
        bv.BaseTypeEngine(device=0, max_samples=16, max_sites=16, n_slots=1)
    lib = capi.load_library()
    ctx = C.c_void_p()
    assert lib.bv_create(0, None, C.byref(ctx)) == -1          # BV_ERR_ARG
    prm = capi.make_params(max_samples=16)
    assert lib.bv_create(0, C.byref(prm), C.byref(ctx)) == -2  # BV_ERR_CUDA
    assert b"no CPU fallback" in lib.bv_last_error(None)


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under basevar_b200/ or include/ may reference it."""
    for d in ("basevar_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, d)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                    txt = open(os.path.join(dirpath, f)).read()
                    assert "oracle" not in txt.lower(), f"{dirpath}/{f} mentions the oracle"


def test_cli_min_af_clamp():
    # std::min(float(100)/n_bam, min_af) in float, widened to double later (src/basetype_caller.cpp:122,506)
    assert bv.cli_min_af(0.01, 1000) == float(np.float32(0.01)) == 0.009999999776482582
    assert bv.cli_min_af(0.01, 100000) == float(np.float32(100.0) / np.float32(100000)) == 0.0010000000474974513
    assert bv.cli_min_af(0.05, 100) == 0.05000000074505806


def test_synth_host_twin_is_deterministic_and_tileable(built_lib):
    m = bv.synth.config_model("C2")
    b, q, s, mq, r = bv.synth_fill_host(m, 100, 64, 1000, with_mapq=True)
    b2, q2, s2, _, r2 = bv.synth_fill_host(m, 132, 32, 1000)
    assert np.array_equal(b[32:], b2) and np.array_equal(q[32:], q2) and np.array_equal(s[32:], s2) and np.array_equal(r[32:], r2)
    assert b.shape == (64, 1008) and (b[:, 1000:] == capi.BASE_N).all() and (s[:, 1000:] == capi.STRAND_NONE).all()
    cov = (b[:, :1000] < 5).mean()
    assert 0.08 < cov < 0.12
    covered = b[:, :1000] < 5
    assert (q[:, :1000][covered] >= 20).all() and (q[:, :1000][covered] <= 40).all() and (q[:, :1000][~covered] == 0).all()
    assert set(np.unique(s[:, :1000][covered])) <= {0, 1} and (s[:, :1000][~covered] == 2).all()
    assert set(np.unique(r)) <= {65, 67, 71, 84}
    assert (mq[:, :1000][covered] >= 10).all()


def test_sparse_host_twin_equals_the_dense_twin(built_lib):
    """bv_synth_fill_sparse_host lists exactly the covered cells of bv_synth_fill_host, packed as BV_CELL_PACK."""
    for cov, N, S in [(0.1, 1000, 200), (0.99326, 7, 300), (0.3, 1003, 50)]:
        m = bv.synth.make_model(seed=5, coverage=cov, variant_frac=0.2, multi_frac=0.2)
        b, q, s, mq, r = bv.synth_fill_host(m, 7, S, N, with_mapq=True)
        rpr = bv.synth_fill_rpr_host(m, 7, S, N)
        cells, aux, st, ref = bv.synth_fill_sparse_host(m, 7, S, N, with_aux=True)
        c2, a2, st2 = bv.dense_to_sparse(b, q, s, N, mq, rpr)
        assert np.array_equal(cells, c2) and np.array_equal(aux, a2) and np.array_equal(st, st2) and np.array_equal(ref, r)
        # unpacking gives the planes back
        site = np.repeat(np.arange(S), np.diff(st.astype(np.int64)))
        samp = cells & ((1 << 20) - 1)
        assert np.array_equal((cells >> 20) & 7, b[site, samp]) and np.array_equal((cells >> 23) & 3, s[site, samp])
        assert np.array_equal(cells >> 25, q[site, samp]) and np.array_equal(aux & 255, mq[site, samp])
        assert np.array_equal(aux >> 8, rpr[site, samp])


def _decode16(words, start, n_samples):
    """Reference decoder of BV_CELLS_U16 (plain Python): list of (site, sample, base, strand, phred, word index)."""
    out = []
    for s in range(len(start) - 1):
        nxt = 0
        for k in range(int(start[s]), int(start[s + 1])):
            w = int(words[k]); gap = w & 31
            if gap == 31:
                assert w == 31
                nxt += 31
                continue
            smp = nxt + gap
            assert smp < n_samples
            out.append((s, smp, (w >> 5) & 7, (w >> 8) & 1, w >> 9, k))
            nxt = smp + 1
    return out


def test_sparse_encode16_round_trip_and_errors(built_lib):
    """The 2-byte delta-coded form decodes to the very cells it was made from; what it cannot carry is refused."""
    rng = np.random.default_rng(3)
    for cov, N, S in [(0.1, 1000, 60), (0.01, 3000, 40), (0.99326, 70, 50), (0.3, 31, 40), (0.05, 63, 40)]:
        m = bv.synth.make_model(seed=11, coverage=cov, variant_frac=0.3, multi_frac=0.3)
        cells, aux, st, _ = bv.synth_fill_sparse_host(m, 3, S, N, with_aux=True)
        words, aux16, st16 = bv.sparse_encode16(cells, st, aux)
        lib = capi.load_library()
        assert len(words) <= lib.bv_sparse_encode16_bound(len(cells), S, N)
        dec = _decode16(words, st16, N)
        assert len(dec) == len(cells)
        site = np.repeat(np.arange(S), np.diff(st.astype(np.int64)))
        for (s, smp, b, sd, q, k), w, a, ws in zip(dec, cells, aux, site):
            assert (s, smp, b, sd, q) == (ws, int(w) & 0xFFFFF, (int(w) >> 20) & 7, (int(w) >> 23) & 3, int(w) >> 25)
            assert aux16[k] == a
    # an empty tile, and sites without cells
    w, _, s16 = bv.sparse_encode16(np.zeros(0, np.uint32), np.zeros(5, np.uint32))
    assert len(w) == 0 and not s16.any()
    # unsorted cells and a strand that is neither + nor - have no compact form
    st = np.array([0, 2], np.uint32)
    with pytest.raises(bv.BvError, match="ascend"):
        bv.sparse_encode16(capi.cell_pack([5, 3], 1, 0, 30), st)
    with pytest.raises(bv.BvError, match="strand"):
        bv.sparse_encode16(capi.cell_pack([3, 5], 1, [0, 2], 30), st)


def test_suggest_tile_sites_without_a_context(built_lib):
    """Multiples of 148 SMs x 32 warps (16 for long rows) that fit the byte budget; small budgets give what fits."""
    lib = capi.load_library()
    assert lib.bv_suggest_tile_sites(None, 1000, 3 * 1008 * 10000) == 2 * 148 * 32
    assert lib.bv_suggest_tile_sites(None, 100000, 3 * 100000 * 10737) == 4 * 148 * 16
    assert lib.bv_suggest_tile_sites(None, 100000, 3 * 100000 * 100) == 100
    assert lib.bv_suggest_tile_sites(None, 1000, 10) == 1
