"""End-to-end (host buffers -> records) throughput of the sparse-tile transport over tile size and slot count.

    python tools/e2e_sweep.py --config C2 --sites 1000000
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import basevar_b200 as bv

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C2")
ap.add_argument("--sites", type=int, default=1000000)
ap.add_argument("--tiles", default="16384,32768,65536,131072")
ap.add_argument("--slots", default="2,3,4")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--u16", action="store_true", help="the compact 2-byte cell form (BV_CELLS_U16)")
args = ap.parse_args()
cfg = dict(bv.synth.CONFIGS[args.config])
N, S = cfg["n_samples"], args.sites
model = bv.synth.make_model(cfg["seed"], cfg["coverage"], cfg["variant_frac"], cfg["multi_frac"])
maf = bv.cli_min_af(0.01, N)
cells, _, site_start, ref = bv.synth_fill_sparse_host(model, 0, S, N, pinned=True)
up_bytes = 4 * cells.shape[0]
if args.u16:
    cells, _, site_start = bv.sparse_encode16(cells, site_start, pinned=True)
    up_bytes = 2 * cells.shape[0]
rec = torch.empty(S * 128, dtype=torch.uint8, pin_memory=True).numpy().view(bv.SITE_OUT_DTYPE)
rec[:] = 0
print(f"{args.config}: {N} samples x {S} sites, {cells.shape[0]} {'u16 words' if args.u16 else 'u32 cells'} ({up_bytes / 1e6:.1f} MB up, {S * 128 / 1e6:.1f} MB down)")
for tile in (int(x) for x in args.tiles.split(",")):
    for slots in (int(x) for x in args.slots.split(",")):
        eng = bv.BaseTypeEngine(device=0, max_samples=N, max_sites=min(tile, S), n_slots=slots, min_af=maf)
        eng.call_sparse(cells, site_start, ref, N, out=rec, out_pinned=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            eng.call_sparse(cells, site_start, ref, N, out=rec, out_pinned=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.reps
        print(f"  tile_sites {tile:7d} slots {slots}: {1e3 * dt:7.2f} ms/step  {S * N / dt / 1e9:7.1f} G sample-sites/s  "
              f"H2D {up_bytes / dt / 1e9:5.1f} GB/s")
        eng.close()
