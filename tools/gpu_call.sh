#!/bin/bash
# Scratch driver for one gpurun call of this round (development only).  Everything lands in gpurun_out/r2/.
out=gpurun_out/r2; mkdir -p $out
tag=${1:-a}
timeout 1800 python -m pytest tests -m gpu -q --durations=8 > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $out/status_$tag.log
tail -15 $out/pytest_$tag.log
timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?" | tee -a $out/status_$tag.log
tail -4 $out/bench_$tag.err
