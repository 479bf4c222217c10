#!/bin/bash
# Scratch driver for one gpurun call of this round (development only).  Everything lands in gpurun_out/r2/.
out=gpurun_out/r2; mkdir -p $out
tag=${1:-a}
for v in "" "CGX_RS_BULK=1"; do
  for cfg in "27 48 0" "27 64 1"; do
    echo "== $v sort_bench $cfg" >> $out/sort_$tag.log
    env $v timeout 180 python tools/sort_bench.py $cfg 3 1 >> $out/sort_$tag.log 2>&1; echo "rc=$?" >> $out/sort_$tag.log
  done
done
cat $out/sort_$tag.log
timeout 1800 python -m pytest tests -m gpu -q --durations=12 > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $out/status_$tag.log
tail -30 $out/pytest_$tag.log
timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?" | tee -a $out/status_$tag.log
tail -6 $out/bench_$tag.err
# where does the reference binary fault on the C1 stand-in?  (one compute-sanitizer tool in this call, no ncu)
python - <<'PY' > $out/sanitize_gen_$tag.log 2>&1
from cgx_b200 import synth
c = synth.generate(10000, 100, v_src=2000, v_tgt=2000, seed=1234, qry_seed=4321)
synth.write_text(c, "/tmp/cgx_san", "corpus")
PY
mkdir -p /tmp/cgx_san/out
( cd /tmp/cgx_san && timeout 900 compute-sanitizer --tool memcheck --print-limit 5 $OLDPWD/oracle/_ref/strmatchcuda corpus.f corpus.q corpus.e corpus.a corpus.lex out ) > $out/sanitize_$tag.log 2>&1
echo "sanitizer rc=$?" | tee -a $out/status_$tag.log
grep -A12 "Invalid\|ERROR SUMMARY" $out/sanitize_$tag.log | head -60
