"""Generate the synthetic BAM fixtures of the BAM-driven packer (run HERE only: needs /root/reference and oracle/_ref).

    bash oracle/build_ref_cli.sh
    gcc -O2 -fPIC -shared -o oracle/_ref/libkeepbf.so oracle/keep_batchfiles.c -ldl
    gcc -O2 -I/root/reference/htslib oracle/bam_index.c oracle/_ref/libhts_ref.a -lz -lm -lpthread -o oracle/_ref/bam_index
    python tests/golden/make_golden_bam.py

What it makes, under tests/golden/bam/ (all of it our own data, none of it from the reference's tests/data):
  ref.fa.gz            two contigs: ctgA (4,000 bp) and ctgB (520,000 bp: longer than the reference's 500-kb creation step)
  sNN.bam, sNN.bam.bai 12 single-sample BAM files written by the BAM writer below and indexed by the reference's own htslib;
                       reads exercise the CIGAR operations M I D N S H = X (padding is left out: the reference's
                       get_aligned_pairs advances the read coordinate over P, bam_record.cpp:251-261, and then reads past the end of
                       the sequence), leading / trailing insertions, deletions
                       next to reference skips, both strands, duplicates, QC failures, unmapped reads with a position,
                       secondary / supplementary reads, mapping qualities around the -q threshold, N bases, and reads
                       crafted onto the calling interval's edges and onto the 500-kb step boundary inside ctgB
  rows.covered.txt.gz  the rows (those with depth > 0, plus the header) of the batchfile the UNMODIFIED reference command
                       wrote for `-r ctgA,ctgB:10001-520000 -q 10 -B 200` (kept with oracle/keep_batchfiles.c)
  syn.vcf.gz syn.cvg.gz, syng.vcf.gz syng.cvg.gz   what the reference wrote, without and with --pop-group
  groups.info          the population groups of the second run
"""
import gzip
import os
import random
import shutil
import struct
import subprocess
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(ROOT, "tests", "golden", "bam")
BIN = os.path.join(ROOT, "oracle", "_ref", "basevar")
KEEP = os.path.join(ROOT, "oracle", "_ref", "libkeepbf.so")
INDEXER = os.path.join(ROOT, "oracle", "_ref", "bam_index")
REGIONS = "ctgA,ctgB:10001-520000"
N_SAMPLES = 12
CONTIGS = (("ctgA", 4000), ("ctgB", 520000))
OPS = "MIDNSHP=X"
SEQ_CODE = {"=": 0, "A": 1, "C": 2, "G": 4, "T": 8, "N": 15}


# ---- BAM writer (SAM specification v1, sections 4.1 and 4.2) ---------------------------------------------------------------
def bgzf_block(data):
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = c.compress(data) + c.flush()
    total = 18 + len(comp) + 8
    return (struct.pack("<4BI2BH2BHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, total - 1) + comp +
            struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data)))


def write_bgzf(path, payload, block=0xff00):
    with open(path, "wb") as f:
        for o in range(0, len(payload), block):
            f.write(bgzf_block(payload[o:o + block]))
        f.write(bgzf_block(b""))


def reg2bin(beg, end):
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def bam_record(name, flag, tid, pos, mapq, cigar, seq, qual):
    """cigar: [(op char, len)], seq: string over ACGTN=, qual: list of phred."""
    ref_len = sum(l for o, l in cigar if o in "MDN=X")
    end = pos + (ref_len if ref_len and not flag & 4 else 1)
    rn = name.encode() + b"\0"
    cig = b"".join(struct.pack("<I", l << 4 | OPS.index(o)) for o, l in cigar)
    codes = [SEQ_CODE[c] for c in seq]
    if len(codes) & 1:
        codes.append(0)
    sq = bytes(codes[i] << 4 | codes[i + 1] for i in range(0, len(codes), 2))
    body = struct.pack("<iiBBHHHiiii", tid, pos, len(rn), mapq, reg2bin(pos, end), len(cigar), flag, len(seq), -1, -1, 0)
    body += rn + cig + sq + bytes(qual)
    return struct.pack("<i", len(body)) + body


def write_bam(path, header_text, reads):
    text = header_text.encode()
    out = [b"BAM\1", struct.pack("<i", len(text)), text, struct.pack("<i", len(CONTIGS))]
    for name, ln in CONTIGS:
        out.append(struct.pack("<i", len(name) + 1) + name.encode() + b"\0" + struct.pack("<i", ln))
    for r in reads:
        out.append(bam_record(*r))
    write_bgzf(path, b"".join(out))


# ---- reads -----------------------------------------------------------------------------------------------------------------
def make_reference(rng):
    seqs = {}
    for name, ln in CONTIGS:
        s = [rng.choice("ACGT") for _ in range(ln)]
        for _ in range(ln // 2000 + 1):        # short runs of N and of soft-masked (lowercase) sequence
            a = rng.randrange(ln)
            for k in range(a, min(ln, a + rng.randrange(1, 8))):
                s[k] = "N"
            a = rng.randrange(ln)
            for k in range(a, min(ln, a + rng.randrange(5, 60))):
                s[k] = s[k].lower()
        seqs[name] = "".join(s)
    return seqs


def random_cigar(rng, rlen):
    """A CIGAR whose reference-consuming operations sum to about rlen."""
    kind = rng.random()
    if kind < 0.55:
        return [("M", rlen)]
    cig = []
    if rng.random() < 0.15:
        cig.append(("H", rng.randrange(1, 6)))
    if rng.random() < 0.30:
        cig.append(("S", rng.randrange(1, 8)))
    if rng.random() < 0.12:
        cig.append(("I", rng.randrange(1, 4)))          # a read that begins with an insertion
    left = rlen
    while left > 0:
        m = min(left, rng.randrange(3, 18))
        cig.append((rng.choice("MMMMM=X"), m))
        left -= m
        if left <= 0:
            break
        x = rng.random()
        if x < 0.30:
            cig.append(("I", rng.randrange(1, 5)))
        elif x < 0.60:
            cig.append(("D", rng.randrange(1, 6)))
        elif x < 0.70:
            cig.append(("N", rng.randrange(5, 60)))
        elif x < 0.75:
            cig.append(("S", 0))                       # a zero-length operation
        elif x < 0.82:
            cig.append(("D", rng.randrange(1, 4)))
            cig.append(("I", rng.randrange(1, 4)))      # an insertion right after a deletion: its anchor is a deleted base
        elif x < 0.88:
            cig.append(("N", rng.randrange(5, 30)))
            cig.append(("I", rng.randrange(1, 3)))
    if rng.random() < 0.10:
        cig.append(("I", rng.randrange(1, 4)))          # a read that ends with an insertion
    if rng.random() < 0.30:
        cig.append(("S", rng.randrange(1, 8)))
    if rng.random() < 0.10:
        cig.append(("H", rng.randrange(1, 6)))
    return cig


def read_for(rng, ref, pos, cigar, snps, name, tid, flag=None, mapq=None):
    seq, qual = [], []
    r = pos
    for op, ln in cigar:
        if op in "M=X":
            for k in range(ln):
                q = rng.randrange(2, 42)
                b = ref[r + k].upper() if r + k < len(ref) else "N"
                if (r + k) in snps and rng.random() < snps[r + k][1]:
                    b = snps[r + k][0]
                if b == "N" or rng.random() < 10 ** (-q / 10.0):
                    b = rng.choice("ACGT")
                if rng.random() < 0.01:
                    b = "N"
                seq.append(b)
                qual.append(q)
            r += ln
        elif op in "IS":
            for _ in range(ln):
                seq.append(rng.choice("ACGT"))
                qual.append(rng.randrange(2, 42))
        elif op in "DN":
            r += ln
    if flag is None:
        flag = 0
        if rng.random() < 0.5:
            flag |= 0x10
        x = rng.random()
        if x < 0.04:
            flag |= 0x400
        elif x < 0.07:
            flag |= 0x200
        elif x < 0.09:
            flag |= 0x4
        elif x < 0.11:
            flag |= 0x100
        elif x < 0.12:
            flag |= 0x800
    if mapq is None:
        mapq = 60 if rng.random() < 0.7 else rng.randrange(0, 60)
    return (name, flag, tid, pos, mapq, cigar, "".join(seq), qual)


def sample_reads(rng, refs, snps, si):
    reads = []
    n = 0

    def add(tid, lo, hi, depth):
        nonlocal n
        ref = refs[CONTIGS[tid][0]]
        count = int((hi - lo) * depth / 40)
        for _ in range(count):
            rlen = rng.randrange(25, 60)
            pos = rng.randrange(max(0, lo - 30), max(1, min(hi, len(ref) - rlen - 80)))
            cig = random_cigar(rng, rlen)
            n += 1
            reads.append(read_for(rng, ref, pos, cig, snps.get(tid, {}), "r%d_%d" % (si, n), tid))

    if si < 10:
        add(0, 0, 4000, 1.5)
        add(1, 9800, 10700, 2.0)
        add(1, 100000, 101200, 2.5)
        add(1, 509700, 510300, 2.5)
        add(1, 519500, 519990, 2.0)
    else:
        # crafted reads only: anchors that no earlier read of the sample occupies (positions below are 0-based starts)
        ref = refs["ctgB"]
        sn = snps.get(1, {})
        crafted = [
            (10000, [("I", 3), ("M", 30)]),                # starts on the interval's first position: anchor 10000 is outside
            (10040, [("S", 4), ("I", 2), ("M", 30)]),      # leading insertion, free anchor inside the interval
            (10100, [("M", 10), ("D", 3), ("I", 2), ("M", 20)]),
            (10200, [("M", 10), ("N", 40), ("I", 2), ("M", 20)]),
            (10300, [("M", 20), ("I", 4)]),                # trailing insertion
            (10400, [("I", 35)]),                          # nothing but an insertion
            (10500, [("S", 35)]),                          # nothing but a soft clip
            (100100, [("M", 30)]),
            (509960, [("M", 30), ("D", 5), ("M", 10)]),    # deletion anchored before the step boundary
            (509985, [("M", 15), ("I", 3), ("M", 20)]),    # 509986..510000, insertion between 510000 | 510001
            (510000, [("I", 3), ("M", 30)]),               # first base 510001: anchor 510000 is the step's last position -> dropped
            (510040, [("D", 4), ("M", 30)]),               # leading deletion
            (519960, [("M", 40)]),                         # ends on the contig's last base
        ] if si == 10 else [
            (9990, [("M", 8), ("D", 4), ("M", 30)]),       # deletion across the interval's first position
            (10060, [("M", 12), ("D", 2), ("N", 10), ("M", 12)]),
            (509990, [("M", 10), ("D", 6), ("M", 20)]),    # 509991..510000 then a deletion of 510001..510006
            (510100, [("H", 5), ("S", 3), ("M", 25), ("N", 7), ("M", 10), ("S", 2)]),
            (519950, [("M", 45), ("I", 2)]),
            (519990, [("M", 10)]),
        ]
        for k, (pos, cig) in enumerate(crafted):
            reads.append(read_for(rng, ref, pos, cig, sn, "c%d_%d" % (si, k), 1, flag=(0x10 if k & 1 else 0), mapq=60))
        add(0, 0, 4000, 0.3)
    reads.sort(key=lambda r: (r[2], r[3]))   # stable: ties keep generation order, which is the file order first-read-wins sees
    return reads


def run_reference(work, tag, extra):
    vcf = os.path.join(work, tag + ".vcf")
    cvg = os.path.join(work, tag + ".cvg")
    shutil.rmtree(os.path.join(work, "cache_" + tag), ignore_errors=True)
    cmd = [BIN, "basetype", "-q", "10", "-B", "200", "-t", "4", "-r", REGIONS, "--output-vcf", vcf, "--output-cvg", cvg,
           "-R", os.path.join(work, "ref.fa"), "-L", os.path.join(work, "bam.list")] + extra
    subprocess.check_call(cmd, env=dict(os.environ, LD_PRELOAD=KEEP), stdout=open(os.path.join(work, tag + ".log"), "w"))
    return vcf, cvg, os.path.join(work, "cache_" + tag)


def main():
    work = os.environ.get("BV_GOLDEN_WORK", "/tmp/bv_golden_bam")
    shutil.rmtree(work, ignore_errors=True)
    os.makedirs(work)
    os.makedirs(OUT, exist_ok=True)
    rng = random.Random(20240817)
    refs = make_reference(rng)
    with open(os.path.join(work, "ref.fa"), "w") as f:
        for name, _ in CONTIGS:
            f.write(">" + name + " synthetic\n")
            s = refs[name]
            for o in range(0, len(s), 60):
                f.write(s[o:o + 60] + "\n")
    # planted SNPs (0-based position -> (alt base, allele frequency)) in the well covered window of ctgB and on ctgA
    snps = {0: {}, 1: {}}
    for tid, lo, hi, k in ((0, 200, 3800, 10), (1, 100050, 101150, 12), (1, 509800, 510200, 6)):
        ref = refs[CONTIGS[tid][0]]
        for _ in range(k):
            p = rng.randrange(lo, hi)
            alt = rng.choice([b for b in "ACGT" if b != ref[p].upper()])
            snps[tid][p] = (alt, rng.choice((0.15, 0.3, 0.5, 0.8, 1.0)))
    bams = []
    for si in range(N_SAMPLES):
        path = os.path.join(work, "s%02d.bam" % si)
        hdr = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % c for c in CONTIGS)
        hdr += "@RG\tID:rg%d\tPL:SYN\tSM:smp%02d\tLB:lib\n" % (si, si)
        if si == 3:
            hdr += "@RG\tID:other\tSM:ignored\n"
        write_bam(path, hdr, sample_reads(rng, refs, snps, si))
        subprocess.check_call([INDEXER, path])
        bams.append(path)
    with open(os.path.join(work, "bam.list"), "w") as f:
        f.write("\n".join(bams) + "\n")
    with open(os.path.join(work, "groups.info"), "w") as f:
        for si in range(N_SAMPLES):
            if si != 7:
                f.write("smp%02d\t%s\n" % (si, "EAS" if si % 3 else "AFR"))
    for tag, extra in (("syn", []), ("syng", ["-G", os.path.join(work, "groups.info")])):
        vcf, cvg, cache = run_reference(work, tag, extra)
        for src, dst in ((vcf, tag + ".vcf.gz"), (cvg, tag + ".cvg.gz")):
            with open(src, "rb") as fi, gzip.GzipFile(os.path.join(OUT, dst), "wb", mtime=0) as fo:
                fo.write(fi.read())
        if tag == "syn":
            rows = []
            for fn in sorted(os.listdir(cache)):
                if not fn.endswith(".bf.gz"):
                    continue
                with gzip.open(os.path.join(cache, fn), "rb") as fi:
                    for line in fi:
                        if line.startswith(b"#") or line.split(b"\t", 4)[3] != b"0":
                            rows.append(line)
            with gzip.GzipFile(os.path.join(OUT, "rows.covered.txt.gz"), "wb", mtime=0) as fo:
                fo.write(b"".join(rows))
            print("covered rows:", sum(1 for r in rows if not r.startswith(b"#")))
        print(tag, "done:", sum(1 for l in open(vcf) if not l.startswith("#")), "VCF records")
    for p in bams:
        shutil.copy(p, OUT)
        shutil.copy(p + ".bai", OUT)
    shutil.copy(os.path.join(work, "groups.info"), OUT)
    with open(os.path.join(work, "ref.fa"), "rb") as fi, gzip.GzipFile(os.path.join(OUT, "ref.fa.gz"), "wb", mtime=0) as fo:
        fo.write(fi.read())


if __name__ == "__main__":
    sys.exit(main())
