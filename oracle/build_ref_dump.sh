#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds oracle/_ref/strmatchcuda_dump: the reference with fwrite hooks
# that dump its intermediate arrays (suffix array, lookup results, gappy hit lists, rule records)
# into $CGX_DUMP_DIR.  The reference sources are copied to a scratch dir under /tmp (never into the
# repo), the hook calls are inserted by line number (the reference snapshot is fixed), and only the
# resulting binary lands in oracle/_ref/ (git-ignored).
set -euo pipefail
REF=${CGX_REFERENCE_DIR:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
[ -d "$REF" ] || { echo "build_ref_dump: $REF not present -- using prebuilt files" >&2; exit 0; }
W=$(mktemp -d /tmp/cgx_refdump.XXXXXX)
cp "$REF"/*.c "$REF"/*.cu "$REF"/*.h "$W"/ && cp -r "$REF/uthash" "$W"/ && chmod -R u+w "$W"
cp "$HERE/ref_dump_hooks.h" "$W"/
# --- SuffixArray.cu: insert from the bottom up so that line numbers stay valid ------------------
sed -i '2233a\
		cgx_dump("twoGapSA", twoGapSA, (size_t)countTwoGapSA*sizeof(twoGapOnSA));\
		cgx_dump("twoGapSearch", twoGapSearch, (size_t)distinctTwoGapCount*sizeof(two_gappy_search));\
		cgx_dump("twogapPattern", qryset->twogapPattern, (size_t)countTwoGapEnu*sizeof(twoGapPattern));' "$W/SuffixArray.cu"
sed -i '1875a\
		cgx_dump("oneGapSA", oneGapSA, (size_t)countOneGapSA*sizeof(oneGapOnSA));\
		cgx_dump("oneGapSearch", oneGapSearch, (size_t)distinctOneGapCount*sizeof(gappy_search));\
		cgx_dump("onegapPattern", qryset->onegapPattern, (size_t)qryset->onegapcount_enu*sizeof(gapPattern));\
		cgx_dump("onegap", qryset->onegap, (size_t)qryset->onegapcount_enu*sizeof(gappy));\
		cgx_dump("precomp_index", ref_h->precomp_index, (size_t)PRECOMPUTECOUNT*PRECOMPUTECOUNT*sizeof(precomp_st_end));\
		cgx_dump("precomp_onegap", ref_h->precomp_onegap, (size_t)ref_h->precomp_count*sizeof(precompute_enu_3));\
		cgx_dump("featureMissingCount", ref_h->featureMissingCount, (size_t)PRECOMPUTECOUNT*PRECOMPUTECOUNT*sizeof(int));\
		cgx_dump("frequentList", ref_h->frequentList, (size_t)PRECOMPUTECOUNT*sizeof(int));' "$W/SuffixArray.cu"
sed -i '1518a\
		cgx_dump("sa", ref_h->sa, (size_t)ref_h->toklen*sizeof(int));\
		cgx_dump("str", ref_h->str, (size_t)ref_h->toklen*sizeof(int));\
		cgx_dump("qrysbuf", qryset->qrysbuf, (size_t)qryset->qrysbufsize);\
		cgx_dump("result_two", qryset->result_two, (size_t)qryset->resbufsize);\
		cgx_dump("connectoffset", qryset->connectoffset, (size_t)qryset->totaltokens*sizeof(int));\
		cgx_dump("result_connect", qryset->result_connect, (size_t)qryset->totalconnect*sizeof(result_t));' "$W/SuffixArray.cu"
sed -i '8a\
#include "ref_dump_hooks.h"' "$W/SuffixArray.cu"
# --- ExtractPair.cu ---------------------------------------------------------------------------
# wall-clock brackets around the three host aggregation calls (createLexiconFast / GappyFast / TwoGapFast, ExtractPair.c:515,664,939 --
# the "ExtractPair.c" CPU leg of SURVEY.md 8d): the reference times its kernels on stderr but not these loops
sed -i '3868a\
	cgx_toc("createLexiconFast", (long)prev_cout);' "$W/ExtractPair.cu"
sed -i '3852a\
	cgx_tic();' "$W/ExtractPair.cu"
sed -i '3792a\
	cgx_toc("createLexiconTwoGapFast", (long)count_two_gap);' "$W/ExtractPair.cu"
sed -i '3765a\
	cgx_tic();' "$W/ExtractPair.cu"
sed -i '3734a\
	cgx_toc("createLexiconGappyFast", (long)count_one_gap);' "$W/ExtractPair.cu"
sed -i '3707a\
	cgx_tic();' "$W/ExtractPair.cu"
sed -i '3672a\
	cgx_dump("out_res", out_res, (size_t)prev_cout*sizeof(res_phrase_t));\
	cgx_dump("oneGapRule", oneGapRule, (size_t)count_one_gap*sizeof(rule_onegap));\
	cgx_dump("twoGapRule", twoGapRule, (size_t)count_two_gap*sizeof(rule_twogap));\
	{ int seps_[4] = {seperatorOneGap, seperatorTwoGap[0], seperatorTwoGap[1], (int)global}; cgx_dump("separators", seps_, sizeof(seps_)); }\
	cgx_dump("blocks", tmp_blocks, (size_t)global*sizeof(saind_t));\
	cgx_dump("RLP", ref_source->RLP, (size_t)ref_source->toklen*sizeof(int));\
	cgx_dump("L_tar", ref_target->L_tar, (size_t)ref_target->toklen);\
	cgx_dump("R_tar", ref_target->R_tar, (size_t)ref_target->toklen);\
	cgx_dump("tgt", ref_target->str, (size_t)ref_target->toklen*sizeof(int));' "$W/ExtractPair.cu"
sed -i '5a\
#include "ref_dump_hooks.h"' "$W/ExtractPair.cu"
# the brackets must sit exactly around the calls (the reference snapshot is fixed; fail loudly if it ever is not)
grep -A1 -n 'cgx_tic();' "$W/ExtractPair.cu" | grep -c 'createLexicon.*Fast(' | grep -qx 3 || { echo "build_ref_dump: timer hooks misplaced" >&2; exit 1; }
grep -B1 -n 'cgx_toc(' "$W/ExtractPair.cu" | grep -c ');' | grep -qx 6 || { echo "build_ref_dump: timer hooks misplaced" >&2; exit 1; }
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}; CXX=${CXX:-g++}
INC="-I$W -I$W/uthash -I/usr/local/cuda/include"
mkdir -p "$W/obj" "$OUT"
pids=()
for f in Disk Main PrintResults SuffixArray Timer ExtractPair; do $CXX -O3 -w -msse4.2 $INC -c "$W/$f.c" -o "$W/obj/$f.o" & pids+=($!); done
for f in Start SuffixArray GappyLook ExtractPair; do $NVCC -gencode arch=compute_100,code=sm_100 -O3 -w -use_fast_math $INC -c "$W/$f.cu" -o "$W/obj/$f.cu_o" & pids+=($!); done
for p in "${pids[@]}"; do wait "$p"; done
$CXX -O3 -o "$OUT/strmatchcuda_dump" "$W"/obj/*.o "$W"/obj/*.cu_o -L/usr/local/cuda/lib64 -lcudart -lm
rm -rf "$W"
echo "build_ref_dump: ok -> $OUT/strmatchcuda_dump"
