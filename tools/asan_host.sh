#!/usr/bin/env bash
# ASan + UBSan build of the C++ host layer and a run of its CPU test programs (no GPU needed):
#   bash tools/asan_host.sh > profiles/r02_asan_ubsan.txt 2>&1
set -u
cd "$(dirname "$0")/.."
O=build/asan; mkdir -p $O
FLAGS="-std=c++17 -O1 -g -fno-omit-frame-pointer -fsanitize=address,undefined -fno-sanitize-recover=undefined -Wall"
HOST="basevar_b200/host/bv_host.cpp basevar_b200/host/bv_caller.cpp basevar_b200/host/bv_bam.cpp basevar_b200/host/bv_pileup.cpp"
echo "== build: g++ $FLAGS"
g++ $FLAGS -fPIC -shared -o $O/libbasevar_b200_host.so $HOST -Lbasevar_b200 -lbasevar_b200 -Wl,-rpath,$PWD/basevar_b200 -lpthread -lz || exit 1
for t in test_host_cpu test_caller_cpu pileup_dump; do
  g++ $FLAGS -o $O/$t tests/cpp/$t.cpp -L$O -lbasevar_b200_host -Lbasevar_b200 -lbasevar_b200 -Wl,-rpath,$PWD/$O -Wl,-rpath,$PWD/basevar_b200 -lpthread -lz -ldl || exit 1
done
export ASAN_OPTIONS=detect_leaks=1:abort_on_error=0 UBSAN_OPTIONS=print_stacktrace=1
rc=0
echo "== test_host_cpu"; $O/test_host_cpu || rc=1
echo "== test_caller_cpu"; $O/test_caller_cpu tests/golden/c1 || rc=1
W=$(mktemp -d); zcat tests/golden/bam/ref.fa.gz > $W/ref.fa; ls $PWD/tests/golden/bam/s*.bam > $W/bam.list
echo "== pileup_dump (12 synthetic BAMs, ctgA + part of ctgB; rows must equal the non-sanitized build's)"
$O/pileup_dump $W/ref.fa $W/bam.list ctgA:1-4000 10 4 | md5sum; tests/cpp/bin/pileup_dump $W/ref.fa $W/bam.list ctgA:1-4000 10 4 | md5sum
$O/pileup_dump $W/ref.fa $W/bam.list ctgB:100001-120000 10 4 333 77 | md5sum; tests/cpp/bin/pileup_dump $W/ref.fa $W/bam.list ctgB:100001-120000 10 4 | md5sum
echo "== pileup_dump (bgzip FASTA, CSI index: the reference's range.bam fixture)"
cp tests/golden/range/range.bam $W/r.bam; cp tests/golden/range/range_csi.bam.csi $W/r.bam.csi; echo $W/r.bam > $W/r.list
$O/pileup_dump tests/golden/range/ce.fa.gz $W/r.list CHROMOSOME_I:900-1200 10 1 | md5sum; zcat tests/golden/range/batch.rows.txt.gz | md5sum
$O/pileup_dump --query-check $W/r.bam 3 300 || rc=1
$O/pileup_dump --bgzf-roundtrip $W/rt.gz 1000003 7 || rc=1
rm -rf $W
echo "== sanitizer run finished, rc=$rc"
exit $rc
