#!/bin/bash
# Hardware counters of the traversal kernels per workload -> gpurun_out/<tag>/trace_counters.json (copied to profiles/ by hand).
# usage (on the GPU box): bash tools/trace_counters.sh <tag> [workload ...]
tag="$1"; shift; out=gpurun_out/$tag; mkdir -p $out
ws="$@"; [ -z "$ws" ] && ws="sponza_1080p classroom_1080p san_miguel_4k"
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__cycles_elapsed.avg,smsp__cycles_active.avg
for w in $ws; do
  PTC_PARTS=1 ncu --metrics $M --clock-control none -k regex:"k_trace|k_shade|k_sky" --csv --log-file $out/tc_$w.csv \
      python tools/render_frames.py --workload $w --frames 3 > $out/tc_$w.log 2>&1
done
python tools/trace_counters.py $out $ws
