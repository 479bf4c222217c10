#!/usr/bin/env bash
# Runs the reference binary (oracle/_ref/strmatchcuda[_dump]) on a synthetic corpus on the GPU box and
# packs its grammar files, stderr log and (dump build) intermediate arrays under gpurun_out/.
# Test infrastructure only.
# usage: tools/run_reference.sh <name> <n_sent> <n_qry> <v> <n_phrases> [dump] [extra generate kwargs]
set -uo pipefail
name=$1; ns=$2; nq=$3; v=$4; np_=$5; mode=${6:-plain}; extra=${7:-}
work=/tmp/cgx_ref_$name
rm -rf $work; mkdir -p $work/out $work/dump gpurun_out
python - <<PY
from cgx_b200 import synth
c = synth.generate($ns, $nq, v_src=$v, v_tgt=$v, n_phrases=$np_ $extra)
synth.write_text(c, "$work", "corpus")
print("generated", c.n_sent, "sentences", len(c.src_words), "tokens", c.n_qry, "queries")
PY
bin=$PWD/oracle/_ref/strmatchcuda
if [ "$mode" = dump ]; then bin=$PWD/oracle/_ref/strmatchcuda_dump; export CGX_DUMP_DIR=$work/dump; fi
( cd $work && $bin corpus.f corpus.q corpus.e corpus.a corpus.lex out > ref_stdout.log 2> ref_stderr.log; echo "exit=$?" >> ref_stderr.log )
tail -3 $work/ref_stderr.log
ls $work/out | wc -l
tar -C $work -czf gpurun_out/ref_$name.tgz out dump ref_stdout.log ref_stderr.log
ls -la gpurun_out/ref_$name.tgz
