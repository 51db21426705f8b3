"""Generate the fixture of the reference's own tiny end-to-end test (run HERE only: needs /root/reference).

    bash oracle/build_ref_cli.sh
    gcc -O2 -fPIC -shared -o oracle/_ref/libkeepbf.so oracle/keep_batchfiles.c -ldl
    gcc -O2 -I/root/reference/htslib oracle/bam_index.c oracle/_ref/libhts_ref.a -lz -lm -lpthread -o oracle/_ref/bam_index
    python tests/golden/make_golden_range.py

The command is /root/reference/tests/data/work.log.sh:1, unchanged:
    basevar basetype --mapq=10 --min-af=0.05 --batch-count=1 --thread=1 --regions=CHROMOSOME_I:900-1200
        --output-vcf vz.vcf --output-cvg t.cvg -R ce.fa.gz -I range.bam -I range.bam
i.e. a BGZF-compressed FASTA (no .fai beside it) and the same BAM given twice.  Committed under tests/golden/range/:
the three input DATA files of that test (ce.fa.gz, range.bam, range.bam.bai: test data, not source code), a CSI index of
the same BAM written by the reference's htslib, and what the unmodified reference wrote: VCF, CVG and the batchfile rows.
"""
import gzip
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("BV_REFERENCE_DIR", "/root/reference")
DATA = os.path.join(REF, "tests", "data")
OUT = os.path.join(ROOT, "tests", "golden", "range")
BIN = os.path.join(ROOT, "oracle", "_ref", "basevar")
KEEP = os.path.join(ROOT, "oracle", "_ref", "libkeepbf.so")
INDEXER = os.path.join(ROOT, "oracle", "_ref", "bam_index")
REGION = "CHROMOSOME_I:900-1200"


def main():
    work = os.environ.get("BV_GOLDEN_WORK", "/tmp/bv_golden_range")
    shutil.rmtree(work, ignore_errors=True)
    os.makedirs(work)
    os.makedirs(OUT, exist_ok=True)
    for f in ("ce.fa.gz", "range.bam", "range.bam.bai"):
        shutil.copy(os.path.join(DATA, f), work)
        os.chmod(os.path.join(work, f), 0o644)
    cmd = [BIN, "basetype", "--mapq=10", "--min-af=0.05", "--batch-count=1", "--thread=1", "--regions=" + REGION,
           "--output-vcf", "vz.vcf", "--output-cvg", "t.cvg", "-R", "ce.fa.gz", "-I", "range.bam", "-I", "range.bam"]
    subprocess.check_call(cmd, cwd=work, env=dict(os.environ, LD_PRELOAD=KEEP), stdout=open(os.path.join(work, "log2"), "w"))
    for src, dst in (("vz.vcf", "vz.vcf.gz"), ("t.cvg", "t.cvg.gz")):
        with open(os.path.join(work, src), "rb") as fi, gzip.GzipFile(os.path.join(OUT, dst), "wb", mtime=0) as fo:
            fo.write(fi.read())
    rows = []
    for d in sorted(os.listdir(work)):
        if not d.startswith("cache_"):
            continue
        for fn in sorted(os.listdir(os.path.join(work, d))):
            if fn.endswith(".bf.gz"):
                rows.append(gzip.open(os.path.join(work, d, fn), "rb").read())
    assert len(rows) == 2 and rows[0] == rows[1]   # --batch-count=1: one batchfile per BAM, and it is the same BAM twice
    with gzip.GzipFile(os.path.join(OUT, "batch.rows.txt.gz"), "wb", mtime=0) as fo:
        fo.write(rows[0])
    for f in ("ce.fa.gz", "range.bam", "range.bam.bai"):
        shutil.copy(os.path.join(DATA, f), OUT)
        os.chmod(os.path.join(OUT, f), 0o644)
    # the same BAM with a CSI index only (min_shift 14): range_csi.bam + range_csi.bam.csi
    shutil.copy(os.path.join(DATA, "range.bam"), os.path.join(work, "range_csi.bam"))
    subprocess.check_call([INDEXER, os.path.join(work, "range_csi.bam"), "14"])
    shutil.copy(os.path.join(work, "range_csi.bam.csi"), OUT)
    n_vcf = sum(1 for l in open(os.path.join(work, "vz.vcf")) if not l.startswith("#"))
    n_cvg = sum(1 for l in open(os.path.join(work, "t.cvg")) if not l.startswith("#"))
    print("range fixture:", n_vcf, "VCF records,", n_cvg, "CVG rows,", rows[0].count(b"\n"), "batchfile lines")


if __name__ == "__main__":
    sys.exit(main())
