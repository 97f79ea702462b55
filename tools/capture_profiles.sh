#!/bin/bash
# Round profile capture (run on the GPU box through gpurun): GPU test log, bench lines of every workload, the ncu launch
# list of the bench command and one `ncu --set full` capture of the hot kernels.  usage: bash tools/capture_profiles.sh <tag>
tag="$1"; out=gpurun_out/$tag; mkdir -p $out
python -m pytest tests -m gpu -q 2>&1 | tail -5 > $out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
python bench.py > $out/bench_sponza_1080p.json 2> $out/bench_sponza_1080p.err
python bench.py --impl reference --steps 4 --warmup 1 > $out/bench_reference_sponza_1080p.json 2>/dev/null
for w in cornell_512:256 dragon_1080p:128 classroom_1080p:128 classroom_1080p_mis:64 san_miguel_4k:32; do
  python bench.py --workload ${w%%:*} --steps ${w##*:} --warmup 4 --cpu-seconds 6 > $out/bench_${w%%:*}.json 2>/dev/null
done
python bench.py --bvh-build lbvh --steps 64 --warmup 4 --no-cpu-baseline > $out/bench_sponza_1080p_lbvh.json 2>/dev/null
python bench.py --workload san_miguel_4k --bvh-build lbvh --steps 16 --warmup 4 --no-cpu-baseline > $out/bench_san_miguel_4k_lbvh.json 2>/dev/null
# launch list of the bench command (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/ncu_launches_bench_steps4.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $out/ncu_launches.log 2>&1
# full capture of the hot kernels, whole-frame launches (PTC_PARTS=1: what the stage-timing pass and the roofline figure use)
PTC_PARTS=1 ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_shade" -s 20 -c 10 -o $out/ncu_full -f \
    python bench.py --steps 2 --warmup 4 --no-cpu-baseline > $out/ncu_full.log 2>&1
ls -la $out
