"""Run the device-resident site kernel a few times on a synthetic config (target for ncu / quick timing).

    python tools/run_kernel.py --config C2 --sites 200000 --launches 3
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import basevar_b200 as bv

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C2")
ap.add_argument("--sites", type=int, default=200000)
ap.add_argument("--samples", type=int, default=0)
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--abs-mode", type=int, default=0)
args = ap.parse_args()
cfg = dict(bv.synth.CONFIGS[args.config])
N = args.samples or cfg["n_samples"]
S = args.sites
pitch = (N + 15) // 16 * 16
dev = torch.device("cuda:0")
maf = bv.cli_min_af(0.01, N)
eng = bv.BaseTypeEngine(device=0, max_samples=N, min_af=maf, abs_mode=args.abs_mode)
eng.synth_set_model(bv.synth.make_model(cfg["seed"], cfg["coverage"], cfg["variant_frac"], cfg["multi_frac"]))
base, qual, strand = (torch.empty((S, pitch), dtype=torch.uint8, device=dev) for _ in range(3))
ref = torch.empty(S, dtype=torch.uint8, device=dev)
out = torch.empty(S * 128, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
eng.synth_fill_device(0, S, N, pitch, base.data_ptr(), qual.data_ptr(), strand.data_ptr(), 0, ref.data_ptr(), st)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.launches + 1)]
ev[0].record()
for i in range(args.launches):
    eng.call_device(base.data_ptr(), qual.data_ptr(), strand.data_ptr(), ref.data_ptr(), S, N, pitch, out.data_ptr(), st)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.launches)]
eng.set_profiling(True)
eng.call_device(base.data_ptr(), qual.data_ptr(), strand.data_ptr(), ref.data_ptr(), S, N, pitch, out.data_ptr(), st)
kt = eng.last_kernel_times()
ekt = eng.last_em_kernel_times()
eng.set_profiling(False)
rec = out.cpu().numpy().view(bv.SITE_OUT_DTYPE)
best = min(ms)
print(f"{args.config} N={N} S={S}: launches ms {['%.3f' % m for m in ms]}; best {S * N / best / 1e6:.1f} G sample-sites/s; "
      f"{S * (3 * N + 128) / best / 1e6:.1f} GB/s algorithmic; variant sites {(rec['n_alt'] > 0).sum()}, "
      f"em_calls mean {rec['em_calls'].mean():.3f}, n_active hist {np.bincount(rec['n_active'], minlength=5).tolist()}; "
      f"kernel ms {' '.join(f'{k[3:-7]}={v:.3f}' for k, v in kt.items())}; bound sites {int(((rec['flags'] & 0x20) != 0).sum())}, "
      f"EM sites {int((rec['em_calls'] >= 3).sum())}; K4 = {' + '.join(f'{k[3:-7]} {v:.3f}' for k, v in ekt.items())}; flags hist {np.bincount(rec['flags'], minlength=256).nonzero()[0].tolist()}")
