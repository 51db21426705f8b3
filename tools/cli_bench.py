"""Time the PRODUCT: `basevar_b200/bin/basevar basetype` against the unmodified reference command on the same box.

    python tools/cli_bench.py --bams 1000 --mb 10 --threads 16              # a generated cohort
    python tools/cli_bench.py --c1 bench_data/c1 --threads 4                # BASELINE.json configs[0] (bam100), when staged

Cohort: `--bams` single-sample BAM files of 100-bp single-end reads over a `--mb` Mb contig at `--depth` x (default 0.1),
reads drawn from a random reference with planted SNPs (allele-frequency spectrum of basevar_b200/synth.py) and phred 20..40
errors, written by the vectorised BAM writer below (SAM specification 4.1 / 4.2) and indexed by the reference's own htslib
(oracle/_ref/bam_index).  Both commands get the same files, region, -q 10 and thread count; ours adds --timing (wall seconds
per stage of the host pipeline).  The VCF and CVG files of the two runs are compared byte for byte (header lines that hold
file paths aside).  One JSON line on stdout.  The reference leg can be skipped (--no-reference) or bounded to a prefix of the
region (--ref-mb): sites are independent and its cost is linear in the region.
"""
import argparse
import gzip
import json
import multiprocessing as mp
import os
import shutil
import struct
import subprocess
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "basevar_b200", "bin", "basevar")
REF = os.path.join(ROOT, "oracle", "_ref", "basevar")
INDEXER = os.path.join(ROOT, "oracle", "_ref", "bam_index")
CONTIG = "chrS"
READ_LEN = 100


def bgzf_blocks(payload, level=1, block=0xff00):
    out = []
    for o in range(0, len(payload), block):
        data = payload[o:o + block]
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        comp = c.compress(data) + c.flush()
        out.append(struct.pack("<4BI2BH2BHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, 18 + len(comp) + 8 - 1) + comp +
                   struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data)))
    out.append(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    return b"".join(out)


def reg2bin(beg, end):
    end = end - 1
    b = np.zeros(beg.shape, np.int64)
    done = np.zeros(beg.shape, bool)
    for shift, off in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        hit = ~done & ((beg >> shift) == (end >> shift))
        b[hit] = off + (beg[hit] >> shift)
        done |= hit
    return b


def make_reference(length, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 4, length, dtype=np.uint8)   # codes 0..3 = A, C, G, T


def site_alleles(length, seed):
    """Planted SNPs: (alt code per site, allele frequency per site; AF 0 = no variant)."""
    rng = np.random.default_rng(seed + 1)
    var = rng.random(length) < 0.019
    u = rng.random(length)
    af = np.where(u < 0.84, 10 ** (-3 + u / 0.84), np.where(u < 0.89, 10 ** (-2 + (u - 0.84) / 0.05 * 0.7),
                  np.where(u < 0.97, 10 ** (-1.3 + (u - 0.89) / 0.08), 0.5 + (u - 0.97) / 0.03 * 0.5)))
    return rng.integers(1, 4, length, dtype=np.uint8), np.where(var, af, 0.0)


def write_sample(job):
    path, sample, length, n_reads, seed, ref, alt_shift, af = job
    rng = np.random.default_rng(seed)
    pos = np.sort(rng.integers(0, length - READ_LEN, n_reads)).astype(np.int64)
    idx = pos[:, None] + np.arange(READ_LEN)[None, :]
    base = ref[idx]
    # this sample's genotype at the read's sites: alt with probability AF (per read: haploid draw, as at < 1x)
    alt = rng.random(base.shape) < af[idx]
    base = np.where(alt, (base + alt_shift[idx]) & 3, base)
    qual = rng.integers(20, 41, base.shape, dtype=np.uint8)
    err = rng.random(base.shape) < 10.0 ** (-qual.astype(np.float64) / 10.0)
    base = np.where(err, (base + rng.integers(1, 4, base.shape, dtype=np.uint8)) & 3, base).astype(np.uint8)
    code = np.array([1, 2, 4, 8], np.uint8)[base]                       # 4-bit codes of A, C, G, T
    seq = (code[:, 0::2] << 4) | code[:, 1::2]
    name_len = 10                                                        # "r%08d" + NUL
    rec_len = 32 + name_len + 4 + READ_LEN // 2 + READ_LEN
    rec = np.zeros((n_reads, 4 + rec_len), np.uint8)
    rec[:, 0:4] = np.frombuffer(struct.pack("<i", rec_len), np.uint8)
    hdr = np.zeros(n_reads, dtype=[("ref", "<i4"), ("pos", "<i4"), ("lname", "u1"), ("mapq", "u1"), ("bin", "<u2"), ("ncig", "<u2"),
                                   ("flag", "<u2"), ("lseq", "<i4"), ("nref", "<i4"), ("npos", "<i4"), ("tlen", "<i4")])
    hdr["pos"] = pos
    hdr["lname"] = name_len
    hdr["mapq"] = np.where(rng.random(n_reads) < 0.9, 60, rng.integers(0, 60, n_reads))
    hdr["bin"] = reg2bin(pos, pos + READ_LEN)
    hdr["ncig"] = 1
    hdr["flag"] = np.where(rng.random(n_reads) < 0.5, 16, 0)
    hdr["lseq"] = READ_LEN
    hdr["nref"] = -1
    hdr["npos"] = -1
    rec[:, 4:36] = hdr.view(np.uint8).reshape(n_reads, 32)
    names = np.char.add("r", np.char.zfill(np.arange(n_reads).astype(str), 8)).astype("S9")
    rec[:, 36:45] = names.view(np.uint8).reshape(n_reads, 9)
    rec[:, 46:50] = np.frombuffer(struct.pack("<I", READ_LEN << 4), np.uint8)   # 100M
    rec[:, 50:50 + READ_LEN // 2] = seq
    rec[:, 50 + READ_LEN // 2:] = qual
    text = ("@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:%s\tLN:%d\n@RG\tID:rg\tPL:SYN\tSM:%s\n" % (CONTIG, length, sample)).encode()
    head = b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", 1) + struct.pack("<i", len(CONTIG) + 1) + CONTIG.encode() + b"\0" + \
        struct.pack("<i", length)
    with open(path, "wb") as f:
        f.write(bgzf_blocks(head + rec.tobytes()))
    subprocess.check_call([INDEXER, path])
    return n_reads


def make_cohort(work, n_bams, length, depth, seed, procs):
    os.makedirs(work, exist_ok=True)
    ref = make_reference(length, seed)
    alt_shift, af = site_alleles(length, seed)
    fa = os.path.join(work, "ref.fa")
    with open(fa, "wb") as f:
        f.write((">%s synthetic\n" % CONTIG).encode())
        letters = np.frombuffer(b"ACGT", np.uint8)[ref].tobytes()
        full = (length // 60) * 60
        rows = np.frombuffer(letters[:full], np.uint8).reshape(-1, 60)
        f.write(np.concatenate([rows, np.full((rows.shape[0], 1), 10, np.uint8)], axis=1).tobytes())
        if length > full:
            f.write(letters[full:] + b"\n")
    with open(fa + ".fai", "w") as f:
        f.write("%s\t%d\t%d\t60\t61\n" % (CONTIG, length, len(">%s synthetic\n" % CONTIG)))
    n_reads = max(1, int(length * depth / READ_LEN))
    jobs = [(os.path.join(work, "s%05d.bam" % i), "smp%05d" % i, length, n_reads, seed * 100003 + i, ref, alt_shift, af) for i in range(n_bams)]
    with mp.Pool(procs) as pool:
        total = sum(pool.imap_unordered(write_sample, jobs, chunksize=4))
    with open(os.path.join(work, "bam.list"), "w") as f:
        f.write("\n".join(j[0] for j in jobs) + "\n")
    return fa, os.path.join(work, "bam.list"), total


def read_text(path):
    return gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()


def body(text):
    return [l for l in text.split("\n") if l and not l.startswith("##contig=") and not l.startswith("##reference=")]


def run(cmd, **kw):
    t0 = time.perf_counter()
    p = subprocess.run(cmd, capture_output=True, text=True, **kw)
    return time.perf_counter() - t0, p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bams", type=int, default=1000)
    ap.add_argument("--mb", type=float, default=10.0)
    ap.add_argument("--depth", type=float, default=0.1)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 4)
    ap.add_argument("--seed", type=int, default=20240011)
    ap.add_argument("--work", default="/tmp/bv_cli_bench")
    ap.add_argument("--c1", default=None, help="directory with the staged bam100 fixture: bam.list (absolute or relative paths) and standin.fa")
    ap.add_argument("--gpus", default="0")
    ap.add_argument("--tile-sites", type=int, default=0)
    ap.add_argument("--no-reference", action="store_true")
    ap.add_argument("--no-ours", action="store_true", help="(no GPU at hand: cohort generation and the reference leg only)")
    ap.add_argument("--ref-mb", type=float, default=0.0, help="time the reference on the first REF-MB Mb only (default: the whole region)")
    ap.add_argument("--keep", action="store_true")
    args = ap.parse_args()

    out = {"threads": args.threads, "gpus": args.gpus}
    if args.c1:
        os.makedirs(args.work, exist_ok=True)
        fa, lst = os.path.abspath(os.path.join(args.c1, "standin.fa")), os.path.join(args.work, "bam.list")
        with open(lst, "w") as f:   # the staged list holds paths relative to the repository root
            f.write("\n".join(l.strip() if os.path.isabs(l.strip()) else os.path.join(ROOT, l.strip())
                              for l in open(os.path.join(args.c1, "bam.list")) if l.strip()) + "\n")
        region, ref_region = "chr11:5246595-5248428,chr17:41197764-41276135", None
        n_bams = sum(1 for l in open(lst) if l.strip())
        out["workload"] = f"C1: the reference's bam100 fixture, {n_bams} BAM files, regions {region} (80,206 positions), -B 200"
        work = args.work
        os.makedirs(work, exist_ok=True)
        n_pos = 80206
    else:
        length = int(args.mb * 1e6)
        work = args.work
        shutil.rmtree(work, ignore_errors=True)
        t0 = time.perf_counter()
        fa, lst, n_reads = make_cohort(work, args.bams, length + 1000, args.depth, args.seed, min(args.threads, os.cpu_count() or 1))
        out["cohort_generation_s"] = round(time.perf_counter() - t0, 1)
        region = f"{CONTIG}:1-{length}"
        ref_len = int(args.ref_mb * 1e6) if args.ref_mb else length
        ref_region = f"{CONTIG}:1-{ref_len}" if ref_len != length else None
        n_bams, n_pos = args.bams, length
        out["workload"] = (f"{args.bams} synthetic single-sample BAM files x {args.mb:g} Mb at {args.depth}x ({n_reads} reads of {READ_LEN} bp), "
                           f"region {region}, -q 10")
    out["sample_sites"] = n_bams * n_pos

    v1, c1 = os.path.join(work, "ours.vcf"), os.path.join(work, "ours.cvg")
    cmd = [OURS, "basetype", "-R", fa, "-L", lst, "-r", region, "-q", "10", "-B", "200", "-t", str(args.threads), "--output-vcf", v1, "--output-cvg", c1,
           "--gpus", args.gpus, "--timing", "--flip-log", os.path.join(work, "ours.flips")]
    if args.tile_sites:
        cmd += ["--tile-sites", str(args.tile_sites)]
    if args.no_ours:
        t_ours, p = 1.0, subprocess.CompletedProcess(cmd, 0, "", "")
        open(v1, "w").close(); open(c1, "w").close()
    else:
        t_ours, p = run(cmd)
    if p.returncode != 0:
        out["error"] = "ours failed: " + (p.stdout + p.stderr)[-1500:]
        print(json.dumps(out))
        return 1
    stage = [l for l in p.stderr.split("\n") if l.startswith("{\"stage_seconds\"")]
    out["ours"] = {"wall_s": round(t_ours, 3), "sample_sites_per_s": n_bams * n_pos / t_ours,
                   "vcf_records": sum(1 for l in open(v1) if not l.startswith("#")), "cvg_rows": sum(1 for l in open(c1) if not l.startswith("#"))}
    if stage:
        out["ours"].update(json.loads(stage[-1]))
    if os.path.exists(os.path.join(work, "ours.flips")):
        out["ours"]["flagged_positions"] = sum(1 for l in open(os.path.join(work, "ours.flips")) if not l.startswith("#"))

    if not args.no_reference and os.path.exists(REF):
        v2, c2 = os.path.join(work, "ref.vcf"), os.path.join(work, "ref.cvg")
        rr = ref_region or region
        cmd = [REF, "basetype", "-R", fa, "-L", lst, "-r", rr, "-q", "10", "-B", "200", "-t", str(args.threads), "--output-vcf", v2, "--output-cvg", c2]
        t_ref, p = run(cmd, cwd=work)
        if p.returncode != 0:
            out["reference"] = {"error": (p.stdout + p.stderr)[-800:]}
        else:
            n_ref_pos = n_pos if not ref_region else int(args.ref_mb * 1e6)
            out["reference"] = {"wall_s": round(t_ref, 3), "region": rr, "sample_sites_per_s": n_bams * n_ref_pos / t_ref,
                                "what": "the unmodified reference command (oracle/_ref/basevar, its own htslib), same files, -q, -B and -t"}
            out["speedup_wall"] = (n_bams * n_pos / t_ours) / (n_bams * n_ref_pos / t_ref)
            if not ref_region:
                out["identical_vcf_cvg"] = bool(body(read_text(v1)) == body(read_text(v2)) and read_text(c1) == read_text(c2))
            else:   # the reference saw a prefix: compare on it
                lim = int(args.ref_mb * 1e6)
                rows = lambda t: [l for l in body(t) if l.startswith("#") or int(l.split("\t")[1]) <= lim]
                out["identical_vcf_cvg_on_reference_region"] = bool(rows(read_text(v1)) == rows(read_text(v2)) and
                                                                   rows(read_text(c1)) == rows(read_text(c2)))
    if not args.keep and not args.c1:
        shutil.rmtree(work, ignore_errors=True)
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
