#!/bin/bash
# Scratch driver for one gpurun call of this round (development only): sort micro-benchmarks of both onesweep forms, the GPU
# test-suite, the reference timers at C1 and a short bench.  Everything lands in gpurun_out/r2/.
out=gpurun_out/r2; mkdir -p $out
tag=${1:-a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi_$tag.log 2>&1
timeout 180 python tools/sort_bench.py 27 48 0 3 1 > $out/sort_bulk_$tag.log 2>&1; rc=$?
echo "sort bulk rc=$rc" | tee -a $out/status_$tag.log
if [ $rc -ne 0 ]; then export CGX_RS_V1=1; echo "falling back to CGX_RS_V1=1 for the rest" | tee -a $out/status_$tag.log; fi
CGX_RS_V1=1 timeout 180 python tools/sort_bench.py 27 48 0 3 0 > $out/sort_v1_$tag.log 2>&1
timeout 180 python tools/sort_bench.py 27 64 1 3 1 > $out/sort_bulk_vals_$tag.log 2>&1; echo "sort bulk vals rc=$?" | tee -a $out/status_$tag.log
CGX_RS_V1=1 timeout 180 python tools/sort_bench.py 27 64 1 3 0 > $out/sort_v1_vals_$tag.log 2>&1
timeout 180 python tools/sort_bench.py 24 23 0 3 1 > $out/sort_bulk_small_$tag.log 2>&1; echo "sort bulk small rc=$?" | tee -a $out/status_$tag.log
timeout 1800 python -m pytest tests -m gpu -q --durations=25 > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $out/status_$tag.log
tail -40 $out/pytest_$tag.log
timeout 600 python tools/ref_timers.py c1 10000 100 2000 > $out/ref_timers_c1_$tag.log 2>&1; echo "ref_timers rc=$?" | tee -a $out/status_$tag.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?" | tee -a $out/status_$tag.log
tail -3 $out/bench_$tag.err
cat $out/sort_bulk_$tag.log $out/sort_v1_$tag.log $out/sort_bulk_vals_$tag.log $out/sort_v1_vals_$tag.log $out/sort_bulk_small_$tag.log
