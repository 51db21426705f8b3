#!/usr/bin/env bash
# Build the UNMODIFIED reference core (BaseType / strand_bias + htslib kfunc.c) into oracle/_ref/.
# Sources are compiled where they lie under /root/reference; nothing is copied into the repo.
# The reference's own CMake cannot be used on Linux, and this path needs only
# basetype.cpp, utils.cpp and htslib/kfunc.c, so the recipe is four compiler calls.
#   libbvref.so         as built by g++: abs() in EM() resolves to int abs(int)   
#   libbvref_dblabs.so  same sources with `-include stdlib.h` => std::abs(double)
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${BV_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
    echo "build_ref.sh: $REF not present; keeping prebuilt files in $OUT" >&2
    exit 0
fi
mkdir -p "$OUT/inc" "$OUT/obj"
: > "$OUT/inc/config.h"    # kfunc.c includes <config.h>; none of its macros matter for kfunc
CXXFLAGS="-std=c++17 -O3 -fPIC"
gcc -O3 -fPIC -c "$REF/htslib/kfunc.c" -I"$OUT/inc" -I"$REF/htslib" -o "$OUT/obj/kfunc.o"
g++ $CXXFLAGS -c "$REF/src/utils.cpp" -o "$OUT/obj/utils.o"
g++ $CXXFLAGS -c "$REF/src/basetype.cpp" -I"$REF/htslib" -o "$OUT/obj/basetype.o"
g++ $CXXFLAGS -include stdlib.h -c "$REF/src/basetype.cpp" -I"$REF/htslib" -o "$OUT/obj/basetype_dblabs.o"
g++ $CXXFLAGS -c "$HERE/ref_shim.cpp" -I"$REF/src" -o "$OUT/obj/ref_shim.o"
g++ $CXXFLAGS -DBVREF_DBLABS -c "$HERE/ref_shim.cpp" -I"$REF/src" -o "$OUT/obj/ref_shim_dblabs.o"
g++ -shared -o "$OUT/libbvref.so" "$OUT/obj/ref_shim.o" "$OUT/obj/basetype.o" "$OUT/obj/utils.o" "$OUT/obj/kfunc.o" -lpthread -lm
g++ -shared -o "$OUT/libbvref_dblabs.so" "$OUT/obj/ref_shim_dblabs.o" "$OUT/obj/basetype_dblabs.o" "$OUT/obj/utils.o" "$OUT/obj/kfunc.o" -lpthread -lm
echo "built $OUT/libbvref.so $OUT/libbvref_dblabs.so"
