#!/bin/bash
# Scratch driver for one gpurun call of this round (development only).  Everything lands in gpurun_out/r2/.
out=gpurun_out/r2; mkdir -p $out
tag=${1:-a}
timeout 2400 python -m pytest tests -m gpu -q --durations=6 > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $out/status_$tag.log
tail -10 $out/pytest_$tag.log
t0=$(date +%s)
timeout 2400 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$? in $(( $(date +%s) - t0 )) s" | tee -a $out/status_$tag.log
tail -4 $out/bench_$tag.err
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?" | tee -a $out/status_$tag.log
bash tools/profile_round.sh r02d > gpurun_out/profile_r02d.log 2>&1; tail -3 gpurun_out/profile_r02d.log
