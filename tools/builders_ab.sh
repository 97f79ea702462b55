for w in sponza_1080p classroom_1080p dragon_1080p san_miguel_4k; do
  for b in sah lbvh ploc; do
    st=32; [ $w = san_miguel_4k ] && st=12
    echo "== $w $b"
    python bench.py --workload $w --bvh-build $b --steps $st --warmup 4 --no-cpu-baseline --no-others --reps 3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('Mrays/s %.1f ms/step %.3f' % (d['value'], d['ms_per_step']), d['config']['bvh'])
"
  done
done
