#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Build the UNMODIFIED reference command line (`basevar`) into oracle/_ref/.
#
# Every source is compiled where it lies under /root/reference (htslib 1.21 vendored there + src/*.cpp); objects and
# the binary go to oracle/_ref/ only; nothing is copied into the repo.  The reference's own build (CMake driving
# htslib's autoconf) is not used: it does not work on Linux.  htslib needs two generated headers,
# config.h and version.h, which this script writes into oracle/_ref/inc (no bz2 / lzma / curl: CRAM blocks in those
# codecs and remote files are irrelevant to the fixtures).
#
# The binary is used (a) by tests/golden/make_golden_cli.py to produce the C1 fixtures (batchfile rows and the VCF / CVG
# text the reference writes for them) and (b) as the end-to-end reference of `bench.py --impl reference --e2e-cli`.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${BV_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
    echo "build_ref_cli.sh: $REF not present; keeping prebuilt files in $OUT" >&2
    exit 0
fi
H="$REF/htslib"
mkdir -p "$OUT/inc_hts" "$OUT/obj/hts"
cat > "$OUT/inc_hts/config.h" <<'EOF'
#define _XOPEN_SOURCE 600
#define HAVE_DRAND48 1
#define HAVE_ATTRIBUTE_CONSTRUCTOR 1
#define HAVE_ATTRIBUTE_TARGET 1
#define HAVE_BUILTIN_CPU_SUPPORT_SSSE3 1
#define HAVE_GETAUXVAL 1
EOF
echo '#define HTS_VERSION_TEXT "1.21"' > "$OUT/inc_hts/version.h"
cat > "$OUT/inc_hts/config_vars.h" <<'EOF2'
#define HTS_CC "gcc"
#define HTS_CPPFLAGS ""
#define HTS_CFLAGS "-O2"
#define HTS_LDFLAGS ""
#define HTS_LIBS "-lz -lm"
EOF2
echo '#define HTSCODECS_VERSION_TEXT "1.6.1"' > "$OUT/inc_hts/version_codecs.h"

HTS_SRC="kfunc kstring bcf_sr_sort bgzf errmod faidx header hfile hts hts_expr hts_os md5 multipart probaln realn regidx
 region sam sam_mods simd synced_bcf_reader vcf_sweep tbx textutils thread_pool vcf vcfutils
 cram/cram_codecs cram/cram_decode cram/cram_encode cram/cram_external cram/cram_index cram/cram_io cram/cram_stats
 cram/mFILE cram/open_trace_file cram/pooled_alloc cram/string_alloc
 htscodecs/htscodecs/arith_dynamic htscodecs/htscodecs/fqzcomp_qual htscodecs/htscodecs/htscodecs
 htscodecs/htscodecs/pack htscodecs/htscodecs/rANS_static4x16pr htscodecs/htscodecs/rANS_static32x16pr_avx2
 htscodecs/htscodecs/rANS_static32x16pr_avx512 htscodecs/htscodecs/rANS_static32x16pr_sse4
 htscodecs/htscodecs/rANS_static32x16pr_neon htscodecs/htscodecs/rANS_static32x16pr htscodecs/htscodecs/rANS_static
 htscodecs/htscodecs/rle htscodecs/htscodecs/tokenise_name3 htscodecs/htscodecs/utils"
CFLAGS="-O2 -fPIC -w -include $OUT/inc_hts/version_codecs.h -I$OUT/inc_hts -I$H -I$H/htscodecs/htscodecs"
OBJS=""
pids=""
for s in $HTS_SRC; do
    o="$OUT/obj/hts/$(echo "$s" | tr '/' '_').o"
    OBJS="$OBJS $o"
    if [ ! -f "$o" ] || [ "$H/$s.c" -nt "$o" ]; then
        extra=""
        case "$s" in
            *avx2) extra="-mavx2 -mpopcnt" ;;
            *avx512) extra="-mavx512f -mpopcnt" ;;
            *sse4) extra="-msse4.1 -mssse3 -mpopcnt" ;;
        esac
        gcc $CFLAGS $extra -c "$H/$s.c" -o "$o" &
        pids="$pids $!"
    fi
done
for p in $pids; do wait "$p"; done
rm -f "$OUT/libhts_ref.a"
ar -rc "$OUT/libhts_ref.a" $OBJS

CXXFLAGS="-std=c++17 -O3 -fPIC -w -I$H"
CPPOBJ=""
pids=""
for f in "$REF"/src/*.cpp; do
    o="$OUT/obj/cli_$(basename "$f" .cpp).o"
    CPPOBJ="$CPPOBJ $o"
    if [ ! -f "$o" ] || [ "$f" -nt "$o" ]; then
        g++ $CXXFLAGS -c "$f" -o "$o" &
        pids="$pids $!"
    fi
done
for p in $pids; do wait "$p"; done
g++ -o "$OUT/basevar" $CPPOBJ "$OUT/libhts_ref.a" -lz -lm -lpthread
echo "built $OUT/basevar"
