"""Stage BASELINE.json configs[0] (the reference's bam100 fixture) where tools/cli_bench.py --c1 and bench.py's `cli` leg find it
on the GPU box: bench_data/c1/ (git-ignored; it travels with the gpurun snapshot).  Run HERE: needs /root/reference.

    python tools/stage_c1.py

Copies the 100 BAM files (+ .bai), writes bam.list with paths relative to the repository root, and the stand-in FASTA of
tests/golden/make_golden_cli.py (true reference bases recovered from CIGAR + MD, 'N' elsewhere; one line per contig) with
its .fai, so that neither command spends its run indexing 216 MB of 'N'.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.golden import make_golden_cli as G   # noqa: E402

OUT = os.path.join(ROOT, "bench_data", "c1")


def main():
    os.makedirs(os.path.join(OUT, "bam100"), exist_ok=True)
    bams = []
    for l in open(os.path.join(G.DATA, "bam100.list")):
        l = l.strip()
        if not l:
            continue
        src = os.path.join(G.DATA, l)
        dst = os.path.join(OUT, "bam100", os.path.basename(l))
        shutil.copy(src, dst)
        for ext in (".bai",):
            if os.path.exists(src + ext):
                shutil.copy(src + ext, dst + ext)
        os.chmod(dst, 0o644)
        bams.append(dst)
    with open(os.path.join(OUT, "bam.list"), "w") as f:
        f.write("\n".join(os.path.relpath(b, ROOT) for b in bams) + "\n")
    fa = os.path.join(OUT, "standin.fa")
    G.build_fasta(bams, fa)
    off = 0
    with open(fa, "rb") as f, open(fa + ".fai", "w") as fai:
        data = f.read()
        while off < len(data):
            e = data.index(b"\n", off)
            name = data[off + 1:e].decode()
            s = e + 1
            e2 = data.index(b"\n", s)
            fai.write("%s\t%d\t%d\t%d\t%d\n" % (name, e2 - s, s, e2 - s, e2 - s + 1))
            off = e2 + 1
    print("staged", len(bams), "BAM files,", os.path.getsize(fa) >> 20, "MiB FASTA ->", OUT)


if __name__ == "__main__":
    main()
