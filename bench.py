#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native basetype core (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2|C3|C5] [--sites S] [--impl ours|reference]

metric  : sample-sites/s of the BaseType core (likelihoods, EM, LRT, QUAL, strand-bias Fisher)
workload: BASELINE.json configs[1] = synthetic 1,000 samples x 1 Mb at 0.1x ("C2"), per GPU.
step    : one pass of the basetype core (kernels K1 count, K2 scalar, K3 bound, K4 EM) over the whole per-GPU workload
          (S sites x N samples).
value   : whole-job sample-sites/s with the planes resident in HBM (kernels only, CUDA events, max over ranks).
e2e     : the same metric through the C ABI with HOST buffers, copies inside the timed region, tiles pipelined over 3
          streams: sparse tiles (pinned u16 words of the covered reads -> bv_tile_submit_sparse: H2D, K0 expand, K1..K4,
          D2H of the 128-byte records -> bv_tile_wait).  e2e_sparse_u32: one u32 per cell; e2e_dense: dense pinned planes
          (bv_tile_submit).
roofline: algorithmic bytes S*(3N+128) per step / average step time (all four kernels), against MEASURED_PEAKS.json
          hbm_gbs; the per-kernel durations are measured live with CUDA events between the kernels (bv_set_profiling).
cpu_baseline: the UNMODIFIED reference (oracle/_ref/libbvref.so: BaseType ctor + lrt() + strand_bias) on the
          host cores, on a bounded prefix of the same workload (rank 0, N=1 only).
Multi-GPU: sites are sharded by contiguous region, one process per GPU, no collective on the data path
          (weak scaling: every GPU gets its own S sites); torch.distributed only for the barrier / max-time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sample-sites/sec, BaseType EM+LRT"
UNIT = "sample-sites/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(config, n_samples, sites):
    """DRAM bytes (read + write) of one step from the committed `ncu --set full` capture of this workload, or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        t = json.load(f)
    e = t.get(config)
    if not e or e.get("n_samples") != n_samples or e.get("sites") != sites:
        return None
    return e["dram_bytes_per_step"]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(local_rank):
    """Run this rank's host thread -- and, by first touch, its pinned staging buffers -- on the CPUs next to its GPU (one
    host worker per GPU, as in the C++ runner).  Returns the CPU list, or None when the topology cannot be read."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local_rank)
        bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [i for i in range(ncpu) if (mask[i // 64] >> (i % 64)) & 1]
        if cpus and len(cpus) < ncpu:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


def reference_arm(args, cfg, n_samples, rank, world):
    """--impl reference: the unmodified reference on the host cores, bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle import loader as L
    import basevar_b200 as bv
    lib = L.load_ref(False)
    cores = os.cpu_count() or 1
    if lib is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libbvref.so was not built (needs /root/reference at build time)"}))
        return
    model = bv.synth.make_model(cfg["seed"], cfg["coverage"], cfg["variant_frac"], cfg["multi_frac"])
    maf = bv.cli_min_af(0.01, n_samples)
    sites = int(min(args.sites or cfg["n_sites"], max(2000, cores * args.ref_sites_per_core)))
    b, q, s, _, r = bv.synth_fill_host(model, 0, sites, n_samples)
    times, cores_t = [], []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        _, core_s = L.ref_tile(b, q, s, r, n_samples, maf, dblabs=bool(args.abs_mode), n_threads=cores)
        wall = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(wall); cores_t.append(core_s)
    # the slowest thread's time inside BaseType/lrt/strand_bias (the shim's BatchInfo fill is not charged)
    t = float(np.mean(cores_t))
    value = sites * n_samples / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.config}: synthetic {n_samples} samples x {cfg['n_sites']} sites at coverage {cfg['coverage']}",
                   "n_samples": n_samples, "sample_sites": sites, "min_af": maf, "em_abs_mode": args.abs_mode},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": f"first {sites} sites x {n_samples} samples; reference BaseType ctor+lrt()+strand_bias per site, "
                                   f"{cores} threads over contiguous site ranges; time = slowest thread inside the reference code "
                                   f"(wall incl. BatchInfo fill {1e3 * float(np.mean(times)):.1f} ms)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=["C2", "C3", "C5"])
    ap.add_argument("--sites", type=int, default=0, help="sites per GPU (default: the config's, capped to fit HBM)")
    ap.add_argument("--abs-mode", type=int, default=0, help="0 = as-built int abs() in EM, 1 = fabs")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--tile-sites", type=int, default=131072, help="sites per host tile of the e2e legs (tools/e2e_sweep.py)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="N > 1: do not bind the rank to the CPUs next to its GPU")
    ap.add_argument("--ref-sites-per-core", type=int, default=40000)
    args = ap.parse_args()

    import basevar_b200 as bv
    from basevar_b200 import shard
    cfg = dict(bv.synth.CONFIGS[args.config])
    n_samples = cfg["n_samples"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, cfg, n_samples, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the basetype core has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 and not args.no_numa_bind else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    pitch = (n_samples + 15) // 16 * 16
    S = args.sites or cfg["n_sites"]
    S = int(min(S, (24 << 30) // (3 * pitch)))  # keep the resident planes <= 24 GB per GPU
    maf = bv.cli_min_af(0.01, n_samples)
    tile_sites = int(min(args.tile_sites, S))
    eng = bv.BaseTypeEngine(device=local_rank, max_samples=n_samples, max_sites=tile_sites, n_slots=3, min_af=maf,
                            abs_mode=args.abs_mode)
    model = bv.synth.make_model(cfg["seed"], cfg["coverage"], cfg["variant_frac"], cfg["multi_frac"])
    eng.synth_set_model(model)

    # ---- resident synthetic planes: this rank's region shard is sites [rank*S, (rank+1)*S) --------------------
    base, qual, strand = (torch.empty((S, pitch), dtype=torch.uint8, device=dev) for _ in range(3))
    ref = torch.empty(S, dtype=torch.uint8, device=dev)
    out = torch.empty(S * 128, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    eng.synth_fill_device(shard.rank_site_range(rank, world, S)[0], S, n_samples, pitch, base.data_ptr(), qual.data_ptr(), strand.data_ptr(), 0,
                          ref.data_ptr(), stream)
    torch.cuda.synchronize()

    def step():
        eng.call_device(base.data_ptr(), qual.data_ptr(), strand.data_ptr(), ref.data_ptr(), S, n_samples, pitch,
                        out.data_ptr(), stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = eng.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    launches = eng.launch_count - launches0
    total_ms_max = shard.max_over_ranks(total_ms, dev)
    value = world * S * n_samples * args.steps / (total_ms_max * 1e-3)

    # ---- per-kernel durations, live (CUDA events between the four kernels, recorded by the library) ------------
    eng.set_profiling(True)
    ksum = None
    for _ in range(args.steps):
        step()
        t_k = eng.last_kernel_times()
        ksum = t_k if ksum is None else {k: ksum[k] + v for k, v in t_k.items()}
    eng.set_profiling(False)
    kernel_ms = {k: v / args.steps for k, v in ksum.items()}

    # ---- end to end through the C ABI with host buffers --------------------------------------------------------
    # (a) sparse tiles (the headline e2e): pinned host arrays of the covered cells (4 bytes each) -> bv_tile_submit_sparse
    #     (H2D of the cells, K0 expand, K1..K4, D2H of the 128-byte records straight into a pinned buffer) -> bv_tile_wait
    t_prep = time.perf_counter()
    site0 = shard.rank_site_range(rank, world, S)[0]
    cells, _, site_start, s_ref = bv.synth_fill_sparse_host(model, site0, S, n_samples, pinned=True)
    t_prep = time.perf_counter() - t_prep
    t1 = time.perf_counter()
    words16, _, start16 = bv.sparse_encode16(cells, site_start, pinned=True)   # 2 bytes per cell, samples delta-coded
    t_prep16 = time.perf_counter() - t1
    rec_sp = torch.empty(S * 128, dtype=torch.uint8, pin_memory=True).numpy().view(bv.SITE_OUT_DTYPE)
    rec_sp[:] = 0

    def sparse_leg(cell_words, starts):
        rec_sp[:] = 0
        eng.call_sparse(cell_words, starts, s_ref, n_samples, out=rec_sp, out_pinned=True)   # warm-up (sizes the cell buffers)
        barrier()
        l0, up0 = eng.launch_count, eng.h2d_bytes
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            eng.call_sparse(cell_words, starts, s_ref, n_samples, out=rec_sp, out_pinned=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        dt_max = shard.max_over_ranks(dt, dev)
        return {"value": world * S * n_samples / dt_max, "ms": 1e3 * dt_max, "launches": eng.launch_count - l0,
                "uploaded": (eng.h2d_bytes - up0) // args.e2e_steps, "records": rec_sp.tobytes()}

    leg32 = sparse_leg(cells, site_start)
    leg16 = sparse_leg(words16, start16)
    sp_launches = leg32["launches"] + leg16["launches"]

    # (b) dense tiles: pinned planes -> bv_tile_submit (H2D of base + strand, qual rows read in place) -> bv_tile_wait
    h_planes = [torch.empty((S, pitch), dtype=torch.uint8, pin_memory=True) for _ in range(3)]
    h_ref = torch.empty(S, dtype=torch.uint8, pin_memory=True)
    for h, d in zip(h_planes, (base, qual, strand)):
        h.copy_(d)
    h_ref.copy_(ref)
    torch.cuda.synchronize()
    hb, hq, hs = (h.numpy() for h in h_planes)
    hr = h_ref.numpy()
    rec = np.zeros(S, dtype=bv.SITE_OUT_DTYPE)
    eng.call_host(hb, hq, hs, hr, n_samples, out=rec)  # warm-up (also first touch of `rec`)
    barrier()
    l0 = eng.launch_count
    up0 = eng.h2d_bytes
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        eng.call_host(hb, hq, hs, hr, n_samples, out=rec)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    e2e_launches = eng.launch_count - l0 + sp_launches
    uploaded = (eng.h2d_bytes - up0) // args.e2e_steps
    e2e_s_max = shard.max_over_ranks(e2e_s, dev)
    e2e_value = world * S * n_samples / e2e_s_max
    clk = clocks.stop()

    # the e2e records (both transports) must be the very records of the device-resident path
    dev_rec = out.cpu().numpy().view(bv.SITE_OUT_DTYPE)
    same = bool(dev_rec.tobytes() == rec.tobytes())
    same_sp = bool(dev_rec.tobytes() == leg16["records"])
    same_sp32 = bool(dev_rec.tobytes() == leg32["records"])

    if rank == 0:
        peak, peak_src = load_peaks()
        avg_ms = float(np.mean(per_launch_ms))
        algo_bytes = S * (3 * n_samples + 128)
        achieved = algo_bytes / (avg_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.config}: synthetic {n_samples} samples x {S} sites per GPU at coverage {cfg['coverage']} "
                                   f"(BASELINE.json configs[1])" if args.config == "C2" else
                                   f"{args.config}: synthetic {n_samples} samples x {S} sites per GPU at coverage {cfg['coverage']}",
                       "n_samples": n_samples, "sites_per_gpu": S, "min_af": maf, "em_abs_mode": args.abs_mode,
                       "variant_sites": int((dev_rec["n_alt"] > 0).sum()), "mean_em_calls": float(dev_rec["em_calls"].mean()),
                       "l2": "inputs (3 planes, %.2f GB) larger than L2; no flush needed" % (3 * S * pitch / 1e9),
                       "parallelism": f"region-sharded x{world}, no collective",
                       "host_binding": (f"rank 0 on the {len(numa_cpus)} CPUs next to its GPU" if numa_cpus else "none")},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": load_traffic(args.config, n_samples, S), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": algo_bytes,
                         "kernel": "the step's four kernels together (K1 bv_count_kernel dominates)", "avg_launch_ms": avg_ms,
                         "kernel_ms": kernel_ms,
                         # K1 alone moves 2 of the 3 planes (base + strand; the qual plane is read only where the result
                         # depends on it): its own bytes / its own time
                         "k1_bytes": S * (2 * n_samples + 128),
                         "k1_achieved": S * (2 * n_samples + 128) / (kernel_ms["bv_count_kernel"] * 1e-3) / 1e9,
                         "k1_frac": S * (2 * n_samples + 128) / (kernel_ms["bv_count_kernel"] * 1e-3) / 1e9 / peak},
            # e2e (headline): sparse host tiles in the compact form.  h2d = 2 bytes per covered cell (+ 4 % "skip" words) + the
            # site offsets + REF bases
            "e2e": {"value": leg16["value"], "unit": UNIT, "h2d_bytes_per_step": int(leg16["uploaded"]), "d2h_bytes_per_step": int(S * 128),
                    "ms_per_step": leg16["ms"], "transport": "sparse tiles, BV_CELLS_U16 (bv_tile_submit_sparse): pinned u16 words of the "
                    "covered reads, sample indices delta-coded; expanded into the dense planes on the device (K0); records DMA'd into "
                    "a pinned buffer", "cells_per_step": int(cells.shape[0]), "words_per_step": int(words16.shape[0]),
                    "tile_sites": tile_sites, "slots": 3, "records_match_device_path": same_sp,
                    "host_prep_s_untimed": t_prep + t_prep16},
            # the same with one self-contained u32 per cell (any cell order within a site)
            "e2e_sparse_u32": {"value": leg32["value"], "unit": UNIT, "h2d_bytes_per_step": int(leg32["uploaded"]),
                               "d2h_bytes_per_step": int(S * 128), "ms_per_step": leg32["ms"], "records_match_device_path": same_sp32},
            # the dense-plane transport of the same workload.  h2d: bytes uploaded by cudaMemcpyAsync (base + strand planes,
            # REF bases) plus the qual rows the kernels read in place from the pinned host plane
            "e2e_dense": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(uploaded + pitch * int((((dev_rec["flags"] & 0x20) != 0) | (dev_rec["em_calls"] >= 3) | ((dev_rec["n_alt"] > 0) & (dev_rec["em_calls"] == 1))).sum())),
                    "uploaded_bytes_per_step": int(uploaded), "d2h_bytes_per_step": int(S * 128),
                    "ms_per_step": 1e3 * e2e_s_max, "tile_sites": tile_sites, "slots": 3, "records_match_device_path": same},
            "gpu_launches": int(launches + e2e_launches),
            "clocks": clk,
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import loader as L
            cores = os.cpu_count() or 1
            if L.load_ref(bool(args.abs_mode)) is not None:
                ns = int(min(S, max(2000, cores * args.ref_sites_per_core)))
                got, core_s = L.ref_tile(hb[:ns], hq[:ns], hs[:ns], hr[:ns], n_samples, maf, dblabs=bool(args.abs_mode), n_threads=cores)
                line["cpu_baseline"] = {"value": ns * n_samples / core_s, "unit": UNIT, "cores": cores, "kind": "reference",
                                        "sample": f"first {ns} sites x {n_samples} samples of the same workload; unmodified reference "
                                                  f"BaseType ctor+lrt()+strand_bias, {cores} threads, slowest thread's time in reference code"}
            else:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": cores, "kind": "reference", "sample": "oracle/_ref not built"}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
