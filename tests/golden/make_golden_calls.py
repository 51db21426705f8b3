#!/usr/bin/env python
"""Generate tests/golden/calls_*.npz: the called-site outputs of the UNMODIFIED reference compiled here (oracle/_ref):
the three rank-sum INFO values (ref_vs_alt_ranksumtest, truncated to int like src/basetype_caller.cpp:1151-1157) and the
population-group calls (BaseType over the group's samples + lrt([REF, ALT...]), src/basetype_caller.cpp:747-797).
Run in the build container only:  python tests/golden/make_golden_calls.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import loader as L  # noqa: E402
from tests import util  # noqa: E402

assert L.ref_available(False) and L.ref_available(True), "build oracle/_ref first (needs /root/reference)"


def main():
    rng = np.random.default_rng(20241017)
    for name, (S, N, cov, qlo, qhi, maf, G, rpr_max) in {
        "calls_N100_g3": (500, 100, 0.5, 0, 40, 0.01, 3, 35),
        "calls_N1000_g5": (300, 1000, 0.1, 2, 41, 0.01, 5, 150),
        "calls_N2000_dense_g2": (60, 2000, 0.99, 2, 41, 0.01, 2, 100),
        "calls_N300_longreads": (120, 300, 0.6, 5, 40, 0.01, 4, 20000),
    }.items():
        b, q, s, r = util.random_tile(rng, S, N, cov, qlo, qhi, other=0.01, indel=0.01)
        # a few sites with lower-case / N reference characters
        r[::17] = np.array([ord(c) for c in "acgtN"], np.uint8)[rng.integers(0, 5, len(r[::17]))]
        mapq, rpr = util.random_aux(rng, b, N, rpr_max)
        grp = util.random_groups(rng, N, G)
        out = {}
        for mode, key in ((0, "int"), (1, "dbl")):
            recs, _ = L.ref_tile(b, q, s, r, N, maf, dblabs=bool(mode))
            calls, groups = L.oracle_calls(b, q, mapq, rpr, r, N, recs, grp, G, maf, mode, use_ref=True)
            out["recs_" + key] = recs.view(np.uint8)
            out["calls_" + key] = calls.view(np.uint8)
            out["groups_" + key] = groups.view(np.uint8).reshape(len(calls), -1)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), base=b, qual=q, strand=s, ref_base=r, mapq=mapq, rpr=rpr,
                            sample_group=grp, n_groups=np.int64(G), n_samples=np.int64(N), min_af=np.float32(maf), **out)
        c = out["calls_int"].view(L.CALL_OUT_DTYPE)
        g = out["groups_int"].view(L.GROUP_OUT_DTYPE).reshape(len(c), G)
        print(f"{name}: {S} sites x {N} samples, {len(c)} called, rank sums != 10000: "
              f"{int((c['mq_rank_sum'] != 10000).sum())}, group calls with ALT: {int((g['n_alt'] > 0).sum())}")


if __name__ == "__main__":
    main()
