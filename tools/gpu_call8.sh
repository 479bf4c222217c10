#!/bin/bash
# One 8-GPU gpurun call: the -g 2 CLI test, the host-ingest ceiling and the bench line at N = 8 (and N = 2).
out=gpurun_out/r2; mkdir -p $out
tag=${1:-a}
nvidia-smi topo -m > $out/topo8_$tag.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -x -k "two_gpus" > $out/pytest8_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $out/status8_$tag.log
tail -3 $out/pytest8_$tag.log
for n in 1 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 tools/d2h_ceiling.py 1800 6 2>/dev/null | tail -1 | tee -a $out/d2h_ceiling_$tag.jsonl
done
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 5 --warmup 3 > $out/bench8_$tag.json 2> $out/bench8_$tag.err; echo "bench8 rc=$?" | tee -a $out/status8_$tag.log
tail -5 $out/bench8_$tag.err
python - <<PY
import json
b=json.load(open("$out/bench8_$tag.json"))
print("N=8 value", b["value"], "e2e", b["e2e"]["value"], "strong", b["strong"]["value"], "bcast", b["index_broadcast"])
print("sweep", [(r["queries"], round(r["value"])) for r in (b.get("sweep") or {}).get("rows", [])])
PY
