"""CPU (-m "not gpu"): the oracle (oracle/bv_oracle.c) is pinned against
  * the literal outputs of the reference binary quoted in SURVEY.md 8c,
  * the committed golden fixtures tests/golden/*.npz (outputs of the compiled, unmodified reference),
  * the compiled reference itself (oracle/_ref) on fresh random tiles, whenever it is present (it is in the build
    container and travels to the GPU box; /root/reference itself is never read here),
  * the known answers of htslib's kf_gammaq / kt_fisher_exact probed from the reference.
Integer fields and call sets must be bit-exact; the floats are compared bit for bit too (same arithmetic, same libm).
"""
import glob
import os

import numpy as np
import pytest

from oracle import loader as L
from tests import util
from tests.golden.sites import GOLDEN_MIN_AF, GOLDEN_SITES, SURVEY_LITERALS

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_FIELDS = ("depth", "depth_other", "fwd", "rev", "n_alt", "alt", "af", "qual", "fs_cvg", "fs_vcf")


def assert_bit_identical(got, want, label):
    for f in REF_FIELDS:
        g, w = got[f], want[f]
        if g.dtype.kind == "f":
            same = (g == w) | (np.isnan(g) & np.isnan(w))
        else:
            same = g == w
        bad = np.nonzero(~same.reshape(len(got), -1).all(axis=1))[0]
        assert len(bad) == 0, (f"{label}: field {f} differs at {len(bad)} sites, first {bad[0]}:\n got  "
                               f"{util.describe(got[bad[0]])}\n want {util.describe(want[bad[0]])}")


def test_survey_literals(oracle_lib):
    b, q, s, r, n = L.planes_from_reads(GOLDEN_SITES)
    for idx, (maf, alts, af_i, qual_i, af_d, qual_d) in SURVEY_LITERALS.items():
        for mode, af, qual in ((0, af_i, qual_i), (1, af_d, qual_d)):
            if af is None:
                continue
            rec = L.oracle_tile(b[idx:idx + 1], q[idx:idx + 1], s[idx:idx + 1], r[idx:idx + 1], n, maf, mode)[0]
            assert "".join("ACGT"[c] for c in rec["alt"][:rec["n_alt"]]) == alts, (idx, mode)
            assert rec["af"][:rec["n_alt"]].tolist() == af, (idx, mode, rec["af"].tolist())
            assert rec["qual"] == qual, (idx, mode, float(rec["qual"]))
    # depths / strand tables of G1, G2, G5 (SB = ref_fwd, ref_rev, alt_fwd, alt_rev)
    rec = L.oracle_tile(b, q, s, r, n, 0.01, 0)
    assert rec["depth"][0].tolist() == [7, 0, 3, 0] and rec["depth"][1].tolist() == [2, 50, 0, 5]
    assert (rec["fwd"][0][0], rec["rev"][0][0], rec["fwd"][0][2], rec["rev"][0][2]) == (3, 4, 1, 2)
    assert (rec["fwd"][1][1], rec["rev"][1][1], rec["fwd"][1][3], rec["rev"][1][3]) == (33, 17, 5, 0)
    assert rec["fs_vcf"][1] == 5.0936618319754317 and rec["fs_vcf"][3] == 3.752293551873672
    assert rec["fs_vcf"][0] == 0 and L.load_oracle().bvo_sor_from_table(3, 4, 1, 2) == 1.5
    assert L.load_oracle().bvo_sor_from_table(20, 20, 35, 25) == 0.7142857142857143
    assert L.load_oracle().bvo_sor_from_table(0, 0, 12, 0) == 10000
    # E1: phred-0 read -> AF NaN, QUAL 0;  E3: indels only -> depth 0
    assert np.isnan(rec["af"][5][0]) and rec["qual"][5] == 0 and rec["n_alt"][5] == 1
    assert rec["depth"][7].sum() == 0 and rec["n_alt"][7] == 0


@pytest.mark.parametrize("path", sorted(p for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                                         if not os.path.basename(p).startswith("calls_")), ids=os.path.basename)
def test_oracle_matches_golden_fixture(oracle_lib, path):
    z = np.load(path)
    if "tables" in z:   # Fisher known answers
        lib = L.load_oracle()
        got = np.array([lib.bvo_fs_from_table(*map(int, t)) for t in z["tables"]])
        assert got.tolist() == z["fs"].tolist()
        return
    n, maf = int(z["n_samples"]), float(z["min_af"])
    for mode, key in ((0, "ref_int"), (1, "ref_dbl")):
        want = z[key].view(L.SITE_OUT_DTYPE)
        got = L.oracle_tile(z["base"], z["qual"], z["strand"], z["ref_base"], n, maf, mode)
        assert_bit_identical(got, want, f"{os.path.basename(path)} mode {mode}")


@pytest.mark.skipif(not L.ref_available(), reason="oracle/_ref not built (needs /root/reference at build time)")
@pytest.mark.parametrize("dbl", [False, True])
def test_oracle_matches_compiled_reference(oracle_lib, dbl):
    if not L.ref_available(dbl):
        pytest.skip("variant not built")
    assert L.load_ref(dbl).bvref_abs_mode() == int(dbl)
    rng = np.random.default_rng(5 + dbl)
    for (S, N, cov, qlo, qhi, maf) in [(1200, 100, 0.5, 0, 40, 0.01), (300, 1000, 0.1, 0, 93, 0.01),
                                       (60, 2000, 0.99, 2, 41, 0.01), (800, 30, 0.4, 0, 93, 0.05),
                                       (40, 10000, 0.1, 20, 40, 0.001)]:
        b, q, s, r = util.random_tile(rng, S, N, cov, qlo, qhi, other=0.01, indel=0.01)
        want, _ = L.ref_tile(b, q, s, r, N, maf, dblabs=dbl, n_threads=2)
        got = L.oracle_tile(b, q, s, r, N, maf, int(dbl))
        assert_bit_identical(got, want, f"N={N} dbl={dbl}")


def test_bad_strand_is_flagged_like_the_reference_throws(oracle_lib):
    """A counted base whose strand is neither + nor -: the reference throws (src/basetype.cpp:271-273)."""
    sites = [("A", [("A", 30, "+"), ("C", 30, "."), ("A", 30, "-")])]
    b, q, s, r, n = L.planes_from_reads(sites)
    rec = L.oracle_tile(b, q, s, r, n, 0.01, 0)
    assert rec["flags"][0] & 0x01
    if L.ref_available():
        want, _ = L.ref_tile(b, q, s, r, n, 0.01)
        assert want["flags"][0] & 0x01


def test_kfunc_known_answers(oracle_lib):
    """kf_gammaq(0.5, chi2/2) and kt_fisher_exact values probed from the reference build."""
    lib = L.load_oracle()
    gq = {0.0: 1.0, 1e-9: 0.99997476867478396, 0.5: 0.47950012218695282, 3.84: 0.050043521248704398,
          23.9: 1.0147176145441744e-06, 24.0: 9.633570086430948e-07, 24.1: 9.1460296797954846e-07,
          100.0: 1.5239706048320983e-23, 1000.0: 1.7958327848007187e-219, 1500.0: 0.0}
    for chi, want in gq.items():
        got = lib.bvo_chi2_test(chi, 1.0)
        assert got == want or abs(got - want) <= 4e-16 * abs(want), (chi, got, want)
    assert np.isnan(lib.bvo_chi2_test(-0.1, 1.0))      # tests/io/test_algorithm.cpp:12 prints nan
    fe = {(0, 0, 0, 0): 1.0, (2, 8, 0, 0): 1.0, (3, 2, 2, 1): 1.0, (4, 3, 3, 0): 0.4749999999999977,
          (500, 480, 20, 3): 0.00050937499055459272, (5000, 4900, 100, 20): 1.0770191917148261e-13,
          (50000, 49000, 300, 200): 2.3839801204085932e-05, (3, 4, 4, 5): 1.0, (1, 1, 1, 1): 1.0}
    for t, want in fe.items():
        got = lib.bvo_fisher_two_sided(*t)
        assert got == want or abs(got - want) <= 1e-15 * abs(want), (t, got, want)
    # the reference's own smoke tables (tests/io/test_algorithm.cpp:20-38), printed with 6 significant digits
    for t, want in {(345, 455, 260, 345): 0.956678, (8, 4, 4, 9): 0.115239, (10, 5, 4, 9): 0.128346}.items():
        assert float(f"{lib.bvo_fisher_two_sided(*t):.6g}") == want


def test_golden_min_af_list():
    assert GOLDEN_MIN_AF == [0.01, 0.05, 0.001]
