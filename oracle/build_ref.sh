#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the *unmodified* reference (hohoCode/cgx) from the
# sources where they lie under /root/reference into oracle/_ref/ (git-ignored, travels to the
# GPU box with gpurun).  Nothing is copied into the repo.  The reference's own Makefile is not
# used (its link line puts the libraries before the objects and fails with a modern ld,
# Makefile:9,31); this recipe compiles each translation unit directly.
#
#   oracle/_ref/strmatchcuda        the reference CLI, compiled for sm_100 (parity pin, GPU box)
#   oracle/_ref/libref_sa.so        SuffixArray.c alone (CPU suffix array / LCP oracle + baseline)
#
# Flags follow the reference Makefile:4-8 (-O3 -use_fast_math, g++ for the .c files) with the
# architecture moved from sm_35 to sm_100.
set -euo pipefail
REF=${CGX_REFERENCE_DIR:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
  echo "build_ref: $REF not present (GPU box?) -- using prebuilt files in $OUT" >&2
  exit 0
fi
mkdir -p "$OUT/obj"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
CXX=${CXX:-g++}
INC="-I$REF -I$REF/uthash -I/usr/local/cuda/include"
pids=()
for f in Disk Main PrintResults SuffixArray Timer ExtractPair; do
  if [ ! -f "$OUT/obj/$f.o" ] || [ "$REF/$f.c" -nt "$OUT/obj/$f.o" ]; then
    $CXX -O3 -w -msse4.2 $INC -c "$REF/$f.c" -o "$OUT/obj/$f.o" &
    pids+=($!)
  fi
done
for f in Start SuffixArray GappyLook ExtractPair; do
  if [ ! -f "$OUT/obj/$f.cu_o" ] || [ "$REF/$f.cu" -nt "$OUT/obj/$f.cu_o" ]; then
    $NVCC -gencode arch=compute_100,code=sm_100 -O3 -w -use_fast_math $INC -c "$REF/$f.cu" -o "$OUT/obj/$f.cu_o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
$CXX -O3 -o "$OUT/strmatchcuda" "$OUT"/obj/*.o "$OUT"/obj/*.cu_o -L/usr/local/cuda/lib64 -lcudart -lm
# CPU-only piece: the reference suffix-array / LCP construction, as a shared object.
$CXX -O3 -w -fPIC -shared $INC "$REF/SuffixArray.c" "$HERE/ref_sa_shim.cpp" -o "$OUT/libref_sa.so"
echo "build_ref: ok -> $OUT/strmatchcuda $OUT/libref_sa.so"
