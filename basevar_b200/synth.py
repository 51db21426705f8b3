"""Synthetic pileup models for the device generator and its host twin.

All probabilities are turned into integer thresholds here, on the host, once; the generator itself
(csrc/bv_synth.cuh) is integer-only, which is what makes host and device bytes identical.
"""
import numpy as np

from .capi import BvSynthModel


def _u32(p):
    return int(min(max(p, 0.0), 1.0) * 4294967296.0) if p < 1.0 else 0xFFFFFFFF


def make_model(seed, coverage=0.1, variant_frac=0.019, multi_frac=0.0, q_lo=20, q_hi=40):
    """coverage: P(cell covered); variant_frac: P(site is variant); multi_frac: P(extra ALT alleles | variant);
    phred uniform in [q_lo, q_hi]."""
    m = BvSynthModel()
    m.seed = seed
    m.cov_thr = min(_u32(coverage), 0xFFFFFFFF)
    m.var_thr = min(_u32(variant_frac), 0xFFFFFFFF)
    m.multi_thr = min(_u32(multi_frac), 0xFFFFFFFF)
    m.q_lo = q_lo
    m.q_span = q_hi - q_lo + 1
    assert 0 <= q_lo <= q_hi <= 93
    for q in range(96):
        m.err_thr[q] = int(10.0 ** (-q / 10.0) * (1 << 24)) if q <= 93 else 0
    # ALT1 allele-frequency spectrum of the chr22 fixture:
    # 84 % log-uniform [1e-3,1e-2), 5 % [1e-2,5e-2), 8 % [5e-2,0.5), 3 % uniform [0.5,1]
    segs = [(0.84, 1e-3, 1e-2, True), (0.05, 1e-2, 5e-2, True), (0.08, 5e-2, 0.5, True), (0.03, 0.5, 1.0, False)]
    for k in range(1024):
        u = (k + 0.5) / 1024.0
        acc = 0.0
        for w, lo, hi, logu in segs:
            if u < acc + w or (w, lo, hi, logu) == segs[-1]:
                t = min(max((u - acc) / w, 0.0), 1.0)
                af = lo * (hi / lo) ** t if logu else lo + (hi - lo) * t
                break
            acc += w
        m.af_thr[k] = min(int(af * 4294967296.0), 0xFFFFFFFF)
    for k in range(256):
        t = (k + 0.5) / 256.0
        af = 0.02 * (0.3 / 0.02) ** t
        m.af_extra_thr[k] = int(af * 4294967296.0)
    return m


# The BASELINE.json configurations.  min_af is what the CLI would use: min(100/N, 0.01).
CONFIGS = {
    "C2": dict(n_samples=1_000, n_sites=1_000_000, seed=20240001, coverage=0.1, variant_frac=0.019, multi_frac=0.0),
    "C3": dict(n_samples=10_000, n_sites=10_000_000, seed=20240002, coverage=0.1, variant_frac=0.019, multi_frac=0.0),
    "C4": dict(n_samples=100_000, n_sites=64_000_000, seed=20240003, coverage=0.1, variant_frac=0.019, multi_frac=0.0),
    "C5": dict(n_samples=2_000, n_sites=1_000_000, seed=20240005, coverage=0.99326, variant_frac=0.5, multi_frac=0.5),
}


def config_model(name):
    c = CONFIGS[name]
    return make_model(c["seed"], c["coverage"], c["variant_frac"], c["multi_frac"])
