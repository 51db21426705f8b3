/* TEST INFRASTRUCTURE ONLY.  `bam_index file.bam [min_shift]` writes file.bam.bai (or, with min_shift > 0, the CSI index
 * file.bam.csi) with the reference's own htslib
 * (sam_index_build, htslib/sam.h), so that the synthetic BAM fixtures of tests/golden/make_golden_bam.py carry the
 * index `samtools index` would have produced.  Built by that script against oracle/_ref/libhts_ref.a:
 *   gcc -O2 -I/root/reference/htslib oracle/bam_index.c oracle/_ref/libhts_ref.a -lz -lm -lpthread -o oracle/_ref/bam_index
 */
#include <stdio.h>
#include <stdlib.h>
#include <htslib/sam.h>

int main(int argc, char** argv) {
    if (argc != 2 && argc != 3) { fprintf(stderr, "usage: bam_index file.bam [min_shift]\n"); return 2; }
    const int rc = sam_index_build(argv[1], argc == 3 ? atoi(argv[2]) : 0);
    if (rc != 0) { fprintf(stderr, "sam_index_build failed: %d\n", rc); return 1; }
    return 0;
}
