#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference compiled here (oracle/_ref, built by
oracle/build_ref.sh from /root/reference).  Run in the build container only:

    python tests/golden/make_golden.py

Each fixture holds the input planes (packed SoA cells, include/basevar_b200.h encoding) and the records the
reference's own BaseType ctor + lrt() + strand_bias produced for them (through oracle/ref_shim.cpp), for the
as-built g++ binary (`ref_int`: int abs() in EM) and for the `-include stdlib.h` variant
(`ref_dbl`: std::abs(double)).  /root/reference does not exist on the GPU box, hence committed vectors.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import loader as L  # noqa: E402
from tests import util  # noqa: E402
from tests.golden.sites import GOLDEN_SITES, GOLDEN_MIN_AF  # noqa: E402

# fields the reference shim fills (n_active / chi2 / em_calls are private to BaseType::lrt)
assert L.ref_available(False) and L.ref_available(True), "build oracle/_ref first (needs /root/reference)"


def ref_both(b, q, s, r, n, maf):
    a, _ = L.ref_tile(b, q, s, r, n, maf, dblabs=False)
    d, _ = L.ref_tile(b, q, s, r, n, maf, dblabs=True)
    return a, d


def save(name, b, q, s, r, n, maf, **extra):
    a, d = ref_both(b, q, s, r, n, maf)
    np.savez_compressed(os.path.join(HERE, name), base=b, qual=q, strand=s, ref_base=r, n_samples=np.int64(n),
                        min_af=np.float32(maf), ref_int=a.view(np.uint8), ref_dbl=d.view(np.uint8), **extra)
    print(f"{name}: {b.shape[0]} sites x {n} samples, min_af {maf!r}, variant sites {(a['n_alt'] > 0).sum()} / "
          f"{(d['n_alt'] > 0).sum()}")


def main():
    # 1. the hand-written vectors of SURVEY.md 8c (G1..G7, E1..E6), at three min_af values
    b, q, s, r, n = L.planes_from_reads(GOLDEN_SITES)
    for maf in GOLDEN_MIN_AF:
        save(f"sites_minaf_{maf}.npz", b, q, s, r, n, maf)
    # 2. random tiles: planted multi-allelic sites, junk characters, indels, phred 0..93
    rng = np.random.default_rng(20240917)
    for name, (S, N, cov, qlo, qhi, maf) in {
        "fuzz_N100": (1500, 100, 0.5, 0, 40, 0.01),
        "fuzz_N1000": (600, 1000, 0.1, 0, 93, 0.01),
        "fuzz_N2000_dense": (120, 2000, 0.99, 2, 41, 0.01),
        "fuzz_N50_minaf05": (1000, 50, 0.3, 0, 93, 0.05),
        "fuzz_N5000_minaf001": (100, 5000, 0.1, 2, 41, 0.001),
    }.items():
        b, q, s, r = util.random_tile(rng, S, N, cov, qlo, qhi, other=0.01, indel=0.01, bad_strand=0.0)
        save(name + ".npz", b, q, s, r, N, maf)
    # 3. known answers of the htslib numerics through the reference's public functions
    lib = L.load_ref(False)
    tables = [(0, 0, 0, 0), (2, 8, 0, 0), (3, 2, 2, 1), (4, 3, 3, 0), (500, 480, 20, 3), (5000, 4900, 100, 20),
              (50000, 49000, 300, 200), (345, 455, 260, 345), (8, 4, 4, 9), (10, 5, 4, 9), (3, 4, 4, 5), (1, 1, 1, 1),
              (33, 17, 5, 0), (20, 20, 35, 25), (600, 0, 0, 400), (11, 22, 33, 44), (0, 5, 7, 0), (1000, 3, 2, 900)]
    fs = np.array([lib.bvref_fisher_fs(*t) for t in tables])
    np.savez_compressed(os.path.join(HERE, "fisher_fs.npz"), tables=np.array(tables, np.int32), fs=fs)
    print("fisher_fs.npz:", fs.tolist())


if __name__ == "__main__":
    main()
