#!/bin/bash
# GPU box: the ncu evidence of one round, reduced to text on the box (the .ncu-rep files are tens of MB each and gpurun
# brings back at most 64 MiB).  Usage: bash tools/profile_round.sh <tag, e.g. r01d>
# Writes gpurun_out/prof_<tag>/: launch list (CSV), per-kernel --set full table (md), hot SASS lines per kernel (txt),
# dram_traffic.json.  Each ncu command runs only after the same command has exited 0 without ncu.
set -u
tag=${1:-r01x}
out=gpurun_out/prof_$tag
mkdir -p $out
python tools/ncu_step.py c2 2 > $out/step_plain.log 2>&1 || { echo "ncu_step failed"; tail -5 $out/step_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/launches_c2.csv python tools/ncu_step.py c2 2 > $out/step_ncu.log 2>&1
cp profiles/dram_traffic.json $out/dram_traffic.json 2>/dev/null
# second batch of the step: lookup, j1, j2, the three extractions, the aggregation of the largest kind (+ the next)
ncu --set full --clock-control none --import-source on -k regex:"j1_pos|j1_scan|j2_ordered|j2_scan|agg_hash|agg_group|agg_rules|extract_|lookup_kernel" --launch-skip 16 --launch-count 13 \
    -o $out/step python tools/ncu_step.py c2 2 > $out/step_full.log 2>&1
python tools/ncu_summary.py $out/step.ncu-rep $out/step_full.md $out/dram_traffic.json > /dev/null 2>&1
# the onesweep pass on its own: 2^27 random 48-bit keys (the shape of the hit sorts), one warm sort skipped
python tools/sort_bench.py 27 48 0 2 0 > $out/sort_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rs_onesweep --launch-skip 6 --launch-count 2 -o $out/sort python tools/sort_bench.py 27 48 0 2 0 > $out/sort_full.log 2>&1
python tools/ncu_summary.py $out/sort.ncu-rep $out/sort_full.md $out/dram_traffic.json > /dev/null 2>&1
# the reports travel back when they fit (64 MiB per call); tools/ncu_sass_hot.py reads them off the box
if [ $(du -sm gpurun_out | cut -f1) -gt 58 ]; then rm -f $out/step.ncu-rep; fi
ls -la $out
