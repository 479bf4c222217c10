#!/bin/bash
out=gpurun_out/r2; mkdir -p $out
tag=${1:-a}
t0=$(date +%s)
timeout 1500 python tools/c4_probe.py 1040000000 50000 1000 200 > $out/c4_probe_$tag.log 2>&1; echo "c4 rc=$? in $(( $(date +%s) - t0 )) s" | tee -a $out/status_$tag.log
tail -16 $out/c4_probe_$tag.log
