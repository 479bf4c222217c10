#!/bin/bash
# Scratch driver for one gpurun call of this round (development only).  Everything lands in gpurun_out/r2/.
out=gpurun_out/r2; mkdir -p $out
tag=${1:-a}
timeout 180 python tools/sort_bench.py 24 23 0 3 1 > $out/sort_bulk_small_$tag.log 2>&1; rc=$?
echo "sort bulk small rc=$rc" | tee -a $out/status_$tag.log
if [ $rc -ne 0 ]; then export CGX_RS_V1=1; echo "falling back to CGX_RS_V1=1 for the rest" | tee -a $out/status_$tag.log; fi
timeout 1800 python -m pytest tests -m gpu -q --durations=25 > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $out/status_$tag.log
tail -45 $out/pytest_$tag.log
timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?" | tee -a $out/status_$tag.log
tail -8 $out/bench_$tag.err
# stall reasons of both onesweep forms (each command exits 0 without ncu right before its ncu run)
timeout 180 python tools/sort_bench.py 27 48 0 2 0 > $out/sort_plain_$tag.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rs_onesweep --launch-skip 6 --launch-count 2 -o $out/sort_bulk_$tag python tools/sort_bench.py 27 48 0 2 0 > $out/sort_ncu_$tag.log 2>&1
CGX_RS_V1=1 timeout 180 python tools/sort_bench.py 27 48 0 2 0 >> $out/sort_plain_$tag.log 2>&1 && \
CGX_RS_V1=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:rs_onesweep --launch-skip 6 --launch-count 2 -o $out/sort_v1_$tag python tools/sort_bench.py 27 48 0 2 0 >> $out/sort_ncu_$tag.log 2>&1
cat $out/sort_plain_$tag.log; ls -la $out
