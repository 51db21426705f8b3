"""FS of 2x2 tables whose two-sided p runs from 1e-280 down into the denormal range: device vs oracle (diagnostic, -m gpu box only)."""
import ctypes, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import basevar_b200 as bv
from oracle import loader

lib = loader.load_oracle()
lib.bvo_fisher_two_sided.restype = ctypes.c_double
base = np.array([19923, 8466, 16895, 13857], float)
tabs = np.array([np.round(base * s) for s in np.linspace(0.80, 1.02, 45)], np.int32)
eng = bv.BaseTypeEngine(device=0, max_samples=int(tabs.sum(axis=1).max()), min_af=0.01)
out = np.zeros(len(tabs))
eng._check(eng.lib.bv_fisher_fs(eng._ctx, tabs.ctypes.data, len(tabs), out.ctypes.data), "bv_fisher_fs")
for t, g in zip(tabs.tolist(), out):
    p = lib.bvo_fisher_two_sided(*t)
    w = -10 * math.log10(p) if p > 0 else 10000.0
    print(t, "oracle p %.4g FS %.6f   device FS %.6f   diff %.3g" % (p, w, g, g - w))
