#!/bin/bash
# Scratch driver for one gpurun call of this round (development only).  Everything lands in gpurun_out/r2/.
out=gpurun_out/r2; mkdir -p $out
tag=${1:-a}
timeout 1800 python -m pytest tests -m gpu -q -x --durations=5 -k "not self_agreement and not c2_full" > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $out/status_$tag.log
tail -8 $out/pytest_$tag.log
timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-c3 --no-gpu-reference > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?" | tee -a $out/status_$tag.log
tail -3 $out/bench_$tag.err
python - <<PY
import json
b=json.load(open("$out/bench_$tag.json"))
print(b['value'],b['ms_per_step'],b['e2e']['value'])
for k,v in b['kernels'].items(): print(k,v['ms_per_step'],v['launches_per_step'],round(v['gbs']))
PY
