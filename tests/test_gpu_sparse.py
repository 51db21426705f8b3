"""-m gpu: sparse host tiles (bv_tile_submit_sparse, kernel K0 bv_expand_kernel) through the C ABI.

The sparse tile is another transport of the same pileup, so the bar is stricter than parity: the records must be
BYTE-identical with those of the dense tile (same kernels on the same planes), and they are checked against the CPU
oracle as well.  Edge cases: empty tiles, sites without cells, a tile with no cell at all, every cell covered, junk
characters / indels / bad strands (any cell that differs from the uncovered triple travels), cells in random order
within a site, malformed input (sample out of range, descending offsets) -> BV_ERR_ARG."""
import ctypes as C

import numpy as np
import pytest

import basevar_b200 as bv
from basevar_b200 import capi
from oracle import loader as L
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine(built_lib):
    eng = bv.BaseTypeEngine(device=0, max_samples=20000, max_sites=1024, n_slots=3, min_af=0.01)
    yield eng
    eng.close()


def _both(engine, b, q, s, r, N, maf, abs_mode, label, shuffle=None):
    engine.set_params(min_af=maf, abs_mode=abs_mode)
    dense = engine.call_host(b, q, s, r, N)
    cells, _, st = bv.dense_to_sparse(b, q, s, N)
    # the compact 2-byte form (ascending samples, strands +/- only) of the same tile
    covered_bad_strand = bool(((s[:, :N] > 1) & ((b[:, :N] != capi.BASE_N) | (q[:, :N] != 0))).any())
    if covered_bad_strand:
        with pytest.raises(bv.BvError, match="strand"):
            bv.sparse_encode16(cells, st)
    else:
        words, _, st16 = bv.sparse_encode16(cells, st)
        assert len(words) >= len(cells)
        compact = engine.call_sparse(words, st16, r, N)
        assert compact.tobytes() == dense.tobytes(), f"{label}: 16-bit sparse and dense records differ"
    if shuffle is not None:   # cell order within a site is free
        for i in range(len(st) - 1):
            shuffle.shuffle(cells[st[i]:st[i + 1]])
    sparse = engine.call_sparse(cells, st, r, N)
    assert sparse.tobytes() == dense.tobytes(), f"{label}: sparse and dense records differ"
    want = L.oracle_tile(b, q, s, r, N, maf, abs_mode)
    ie, fe, flips = util.compare_records(sparse, want)
    assert len(ie) == 0 and len(fe) == 0, f"{label}: {len(ie)} exact / {len(fe)} float mismatches vs the oracle"
    assert len(flips) <= 2
    return sparse


@pytest.mark.parametrize("abs_mode", [0, 1])
@pytest.mark.parametrize(
    "name,S,N,kw",
    [
        ("C2-like", 5000, 1000, dict(coverage=0.1, variant_frac=0.05)),
        ("C1-like-N100", 3000, 100, dict(coverage=0.065, variant_frac=0.05)),
        ("C3-like", 1100, 10000, dict(coverage=0.1, variant_frac=0.1)),            # 10 staged chunks per row
        ("long-rows", 300, 20000, dict(coverage=0.05, variant_frac=0.2)),          # K0's direct path (pitch > 16 KB)
        ("chunk-edge", 1500, 1024, dict(coverage=0.2, variant_frac=0.2)),          # row == one staged chunk
        ("chunk-edge+1", 1500, 1025, dict(coverage=0.2, variant_frac=0.2)),        # a 16-byte second chunk
        ("C5-like", 1100, 2000, dict(coverage=0.99326, variant_frac=0.5, multi_frac=0.5)),
        ("odd-N", 2500, 1003, dict(coverage=0.3, variant_frac=0.3, multi_frac=0.5)),
        ("tiny-N", 2500, 7, dict(coverage=0.8, variant_frac=0.5)),
    ],
)
def test_sparse_equals_dense_and_oracle(engine, name, S, N, kw, abs_mode):
    model = bv.synth.make_model(seed=4321 + N, **kw)
    b, q, s, _, r = bv.synth_fill_host(model, 0, S, N)
    got = _both(engine, b, q, s, r, N, bv.cli_min_af(0.01, N), abs_mode, name)
    assert (got["n_alt"] > 0).sum() > 0


def test_sparse_rows_of_100000_samples(built_lib):
    """BASELINE configs[3] rows (100,000 samples, 98 staged chunks per row) through both cell formats: the 16-bit words are consumed
    front to back across the chunks (stream_cells16), the 32-bit cells take K0's direct path."""
    N, S = 100_000, 48
    model = bv.synth.make_model(seed=77, coverage=0.1, variant_frac=0.3)
    b, q, s, _, r = bv.synth_fill_host(model, 5_000_000, S, N)
    eng = bv.BaseTypeEngine(device=0, max_samples=N, max_sites=32, n_slots=2, min_af=bv.cli_min_af(0.01, N))
    try:
        _both(eng, b, q, s, r, N, bv.cli_min_af(0.01, N), 0, "N=100000")
        # a site whose only cells sit at the very end of the row, and one with a single cell at sample 0
        b2 = np.full_like(b[:3], 5); q2 = np.zeros_like(q[:3]); s2 = np.full_like(s[:3], 2)
        b2[0, N - 3:N] = [0, 1, 2]; q2[0, N - 3:N] = 30; s2[0, N - 3:N] = [0, 1, 0]
        b2[1, 0] = 3; q2[1, 0] = 40; s2[1, 0] = 1
        _both(eng, b2, q2, s2, np.array([65, 67, 71], np.uint8), N, bv.cli_min_af(0.01, N), 0, "N=100000 sparse ends")
    finally:
        eng.close()


def test_sparse_generator_twin(engine):
    """bv_synth_fill_sparse_host == the dense host twin in sparse form, and runs to the same records."""
    model = bv.synth.config_model("C2")
    N, S = 1000, 3000
    b, q, s, _, r = bv.synth_fill_host(model, 123456, S, N)
    cells, _, st, ref = bv.synth_fill_sparse_host(model, 123456, S, N, pinned=True)
    c2, _, st2 = bv.dense_to_sparse(b, q, s, N)
    assert np.array_equal(cells, c2) and np.array_equal(st, st2) and np.array_equal(ref, r)
    engine.set_params(min_af=bv.cli_min_af(0.01, N), abs_mode=0)
    import torch
    out = torch.empty(S * 128, dtype=torch.uint8, pin_memory=True).numpy().view(bv.SITE_OUT_DTYPE)
    out[:] = 0
    got = engine.call_sparse(cells, st, ref, N, out=out, out_pinned=True)     # records land by DMA in the pinned buffer
    assert got.tobytes() == engine.call_host(b, q, s, r, N).tobytes()


def test_sparse_fuzz_junk_and_random_cell_order(engine):
    rng = np.random.default_rng(11)
    for (S, N, cov, qlo, qhi, maf) in [(1500, 100, 0.5, 0, 40, 0.01), (1200, 1000, 0.1, 0, 93, 0.01), (600, 2000, 0.99, 2, 41, 0.01)]:
        b, q, s, r = util.random_tile(rng, S, N, cov, qlo, qhi, other=0.01, indel=0.01, bad_strand=0.001)
        _both(engine, b, q, s, r, N, maf, 0, f"fuzz N={N}", shuffle=rng)


def test_sparse_edge_cases(engine):
    engine.set_params(min_af=0.01, abs_mode=0)
    N = 50
    # no site at all
    got = engine.call_sparse(np.zeros(0, np.uint32), np.zeros(1, np.uint32), np.zeros(0, np.uint8), N)
    assert got.shape == (0,)
    # sites without a single cell: all-zero records, like the dense path
    S = 1500   # more than one tile
    ref = np.full(S, ord("A"), np.uint8)
    got = engine.call_sparse(np.zeros(0, np.uint32), np.zeros(S + 1, np.uint32), ref, N)
    b = np.full((S, 64), 5, np.uint8); q = np.zeros((S, 64), np.uint8); s = np.full((S, 64), 2, np.uint8)
    assert got.tobytes() == engine.call_host(b, q, s, ref, N).tobytes()
    assert not got["depth"].any()
    # every cell covered, one site empty in the middle
    rng = np.random.default_rng(5)
    b, q, s, r = util.random_tile(rng, 300, N, 1.0, 20, 40)
    b[17, :] = 5; q[17, :] = 0; s[17, :] = 2
    _both(engine, b, q, s, r, N, 0.01, 0, "full coverage")


def test_sparse_malformed_input_is_an_error(engine):
    engine.set_params(min_af=0.01, abs_mode=0)
    N, S = 50, 10
    ref = np.full(S, ord("C"), np.uint8)
    cells = capi.cell_pack(np.arange(S) % N, 1, 0, 30)
    st = np.arange(S + 1, dtype=np.uint32)
    engine.call_sparse(cells, st, ref, N)                      # well formed
    bad = cells.copy(); bad[3] = capi.cell_pack(N, 1, 0, 30)   # sample == n_samples
    with pytest.raises(bv.BvError, match="sparse tile"):
        engine.call_sparse(bad, st, ref, N)
    st2 = st.copy(); st2[4], st2[5] = st[5], st[4]             # descending offsets
    with pytest.raises(bv.BvError, match="sparse tile"):
        engine.call_sparse(cells, st2, ref, N)
    engine.call_sparse(cells, st, ref, N)                      # the slot is usable again
    # compact form: a gap that runs past the last sample
    words, _, st16 = bv.sparse_encode16(cells, st)
    engine.call_sparse(words, st16, ref, N)
    w2 = words.copy(); w2[7] = (w2[7] & ~np.uint16(31)) | np.uint16(30)   # site 7's only cell: sample 7 -> 30 ... fine; then skips
    long_ = np.concatenate([w2[:8], np.full(2, 31, np.uint16), w2[7:8], w2[8:]])   # site 7: cell, skip, skip, cell at 30+1+62+30 >= 50
    st3 = st16.copy(); st3[8:] += 3
    with pytest.raises(bv.BvError, match="sparse tile"):
        engine.call_sparse(long_, st3, ref, N)
    engine.call_sparse(words, st16, ref, N)


@pytest.mark.parametrize("G", [0, 3])
def test_sparse_calls_equal_dense_calls(built_lib, G):
    """Called-site kernels on a sparse tile (mapq / rpr travel as the cells' aux words)."""
    eng = bv.BaseTypeEngine(device=0, max_samples=2000, max_sites=512, n_slots=2, min_af=0.01)
    try:
        rng = np.random.default_rng(21 + G)
        for (S, N, kw) in [(1500, 1000, dict(coverage=0.1, variant_frac=0.2)), (700, 2000, dict(coverage=0.99326, variant_frac=0.5, multi_frac=0.5))]:
            model = bv.synth.make_model(seed=99 + N, **kw)
            b, q, s, mapq, r = bv.synth_fill_host(model, 0, S, N, with_mapq=True)
            rpr = bv.synth_fill_rpr_host(model, 0, S, N)
            eng.set_params(min_af=bv.cli_min_af(0.01, N), abs_mode=0)
            eng.set_groups(util.random_groups(rng, N, G) if G else None, G)
            d_rec, d_calls, d_groups = eng.call_host_calls(b, q, s, r, mapq, rpr, N)
            cells, aux, st, ref = bv.synth_fill_sparse_host(model, 0, S, N, with_aux=True)
            s_rec, s_calls, s_groups = eng.call_sparse_calls(cells, aux, st, ref, N)
            assert s_rec.tobytes() == d_rec.tobytes()
            assert len(s_calls) > 0 and s_calls.tobytes() == d_calls.tobytes()
            assert s_groups.tobytes() == d_groups.tobytes()
            words, aux16, st16 = bv.sparse_encode16(cells, st, aux)          # the compact form with re-indexed aux words
            c_rec, c_calls, c_groups = eng.call_sparse_calls(words, aux16, st16, ref, N)
            assert c_rec.tobytes() == d_rec.tobytes() and c_calls.tobytes() == d_calls.tobytes() and c_groups.tobytes() == d_groups.tobytes()
    finally:
        eng.close()


@pytest.mark.parametrize("fmt16", [False, True])
def test_compact_record_transport(fmt16):
    """BV_OUT_COMPACT: 8 bytes per site plus a full record for the sites that need one; expanded, the records are byte for byte
    those of the ordinary transport, and bv_tile_wait() of a compact tile expands them itself."""
    N, S = 1000, 20000
    model = bv.synth.make_model(seed=777, coverage=0.1, variant_frac=0.05, multi_frac=0.3)
    maf = bv.cli_min_af(0.01, N)
    cells, _, st, ref = bv.synth_fill_sparse_host(model, 5000, S, N)
    if fmt16:
        cells, _, st = bv.sparse_encode16(cells, st)
    eng = bv.BaseTypeEngine(device=0, max_samples=N, max_sites=4096, n_slots=3, min_af=maf)
    try:
        want = eng.call_sparse(cells, st, ref, N)
        got = np.zeros(S, capi.SITE_OUT_DTYPE)
        stats = {"full": 0, "tiles": 0}

        def take(s0, ns, briefs, full):
            stats["full"] += len(full); stats["tiles"] += 1
            assert len(briefs) == ns
            eng.expand_compact(briefs, full, ref[s0:s0 + ns], got[s0:s0 + ns])

        tiles = eng.sparse_tiles(cells, st, ref, N, compact=True)
        eng.run_sparse_tiles(tiles, 2, compact=take)
        assert got.tobytes() == want.tobytes()
        assert stats["tiles"] == 2 * len(tiles) and 0 < stats["full"] < 2 * 0.25 * S      # ~12 % of the sites need a full record
        # a compact tile collected with bv_tile_wait(): the library expands it
        t = tiles[1][0]
        eng._check(eng.lib.bv_tile_submit_sparse(eng._ctx, 0, C.byref(t)), "submit")
        out = np.zeros(int(t.n_sites), capi.SITE_OUT_DTYPE)
        eng._check(eng.lib.bv_tile_wait(eng._ctx, 0, out.ctypes.data), "wait")
        assert out.tobytes() == want[4096:4096 + int(t.n_sites)].tobytes()
    finally:
        eng.close()
