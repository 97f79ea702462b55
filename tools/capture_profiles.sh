#!/bin/bash
# Round profile capture (run on the GPU box through gpurun): GPU test log, bench lines of every workload, the ncu launch
# list of the bench command, one `ncu --set full` capture of the hot kernels and the traversal counters behind
# roofline.issue.  usage: bash tools/capture_profiles.sh <tag> [quick]
tag="$1"; quick="$2"; out=gpurun_out/$tag; mkdir -p $out
python -m pytest tests -m gpu -q 2>&1 | tail -5 > $out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
# hardware counters of the traversal kernels first: bench.py reads profiles/trace_counters.json for roofline.issue
bash tools/trace_counters.sh $tag sponza_1080p classroom_1080p san_miguel_4k > $out/trace_counters.log 2>&1
cp $out/trace_counters.json profiles/trace_counters.json
python bench.py > $out/bench_sponza_1080p.json 2> $out/bench_sponza_1080p.err
python bench.py --impl reference --steps 4 --warmup 1 > $out/bench_reference_sponza_1080p.json 2>/dev/null
if [ -z "$quick" ]; then
  for w in cornell_512:256 dragon_1080p:128; do
    python bench.py --workload ${w%%:*} --steps ${w##*:} --warmup 4 --cpu-seconds 6 > $out/bench_${w%%:*}.json 2>/dev/null
  done
  for b in lbvh ploc; do
    python bench.py --bvh-build $b --steps 64 --warmup 4 --no-cpu-baseline --no-others > $out/bench_sponza_1080p_$b.json 2>/dev/null
    python bench.py --workload san_miguel_4k --bvh-build $b --steps 16 --warmup 4 --no-cpu-baseline > $out/bench_san_miguel_4k_$b.json 2>/dev/null
  done
fi
# launch list of the bench command (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/ncu_launches_bench_steps4.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-others > $out/ncu_launches.log 2>&1
# full capture of the hot kernels, whole-frame launches (PTC_PARTS=1: what the stage-timing pass and the roofline figure use)
PTC_PARTS=1 ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_shade|k_sky" -s 22 -c 11 -o $out/ncu_full -f \
    python bench.py --steps 2 --warmup 4 --no-cpu-baseline --no-others > $out/ncu_full.log 2>&1
ls -la $out
