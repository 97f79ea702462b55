# usage: bash tools/ab_run.sh <variant|cur[:ENV=VAL,...]>...   A/B bench of library variants in one gpurun call
for spec in "$@"; do
  v="${spec%%:*}"; envs=""; [ "$spec" != "$v" ] && envs="${spec#*:}"
  if [ "$v" != "cur" ]; then export PTC_LIB=$PWD/build/$v/pkg/csrc/libptcore.so; else unset PTC_LIB; fi
  echo "== variant '$spec'"
  env $(echo "$envs" | tr ',' ' ') timeout 300 python bench.py --steps ${STEPS:-32} --warmup 4 --no-cpu-baseline --no-others --reps 3 ${BENCH_ARGS} 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'].get('bvh')); print('Mrays/s %.1f  e2e %.1f  ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), {k: round(v,2) for k,v in d['roofline']['stage_ms'].items()})
    else: print(l.rstrip())
"
done
