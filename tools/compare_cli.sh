#!/usr/bin/env bash
# GPU box: run the reference binary, the product CLI and the CPU oracle CLI on the same synthetic
# corpus and compare the grammar directories as multisets (cgx_b200/grammar_compare.py).
# usage: tools/compare_cli.sh <name> <n_sent> <n_qry> <v> <n_phrases> [extra generate kwargs]
set -uo pipefail
name=$1; ns=$2; nq=$3; v=$4; np_=$5; extra=${6:-}
work=/tmp/cgx_cmp_$name
rm -rf $work; mkdir -p $work/ref $work/mine $work/orc gpurun_out
make -s -C oracle >/dev/null 2>&1
python - <<PY
from cgx_b200 import synth
c = synth.generate($ns, $nq, v_src=$v, v_tgt=$v, n_phrases=$np_ $extra)
synth.write_text(c, "$work", "corpus")
print("generated", c.n_sent, "sentences", len(c.src_words), "tokens", c.n_qry, "queries")
PY
here=$PWD
cd $work
args="corpus.f corpus.q corpus.e corpus.a corpus.lex"
( time $here/oracle/_ref/strmatchcuda $args ref > ref.out 2> ref.err ) 2>&1 | grep real | sed 's/^/reference /'
( time $here/bin/strmatchcuda -q $args mine > mine.out 2> mine.err ) 2>&1 | grep real | sed 's/^/product   /'
tail -2 mine.err
( time $here/oracle/_build/cgx_oracle_cli $args orc > orc.out 2> orc.err ) 2>&1 | grep real | sed 's/^/oracle    /'
cd $here
echo "== product vs oracle"; python -m cgx_b200.grammar_compare $work/mine $work/orc | python -c "import json,sys; d=json.load(sys.stdin); d.pop('examples'); print(d)"
echo "== product vs reference"; python -m cgx_b200.grammar_compare $work/mine $work/ref | python -c "import json,sys; d=json.load(sys.stdin); d.pop('examples'); print(d)"
echo "== oracle vs reference"; python -m cgx_b200.grammar_compare $work/orc $work/ref | python -c "import json,sys; d=json.load(sys.stdin); d.pop('examples'); print(d)"
cmp <(cat $work/mine/grammar.0.s) <(cat $work/orc/grammar.0.s) && echo "grammar.0.s byte-identical product/oracle"
