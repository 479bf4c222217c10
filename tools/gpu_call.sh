#!/bin/bash
out=gpurun_out/r2; mkdir -p $out
tag=${1:-a}
timeout 1800 python -m pytest tests -m gpu -q -x -k "not self_agreement and not c2_full and not c1_grammar" > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $out/status_$tag.log
tail -4 $out/pytest_$tag.log
timeout 1500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-sweep > $out/bench_$tag.json 2> $out/bench_$tag.err
python - <<PY
import json
b=json.load(open("$out/bench_$tag.json"))
k=b['kernels']
print("step %.2f ms, e2e %.0f" % (b['ms_per_step'], b['e2e']['value']), {n: round(v['ms_per_step'],2) for n,v in k.items() if v['ms_per_step']>2})
c=b['c3']; print('c3', round(c['value']), [round(x['ms_total']) for x in c['per_batch']][:4], {n: round(v['ms']/20,1) for n,v in c['kernels'].items() if v['ms']>400})
PY
