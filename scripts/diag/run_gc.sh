for gc in 16,8 8,4 12,4 4,2 8,8 4,4; do
  ACEQD_SPLITK_GC=$gc ACEQD_KERNEL=splitk python bench.py --workload cfg3 --steps 2 --warmup 1 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('cfg3 GC=$gc', 'k_ms', round(r['kernel_ms'],2), 'frac', round(r['frac'],3), 'T', r['tile_T'])"
done
for gc in 8,4 4,4 6,8 7,8 3,8 2,4; do
  ACEQD_SPLITK_GC=$gc python scripts/bench_shapes.py --kernel=splitk 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    try:
        x=json.loads(l); print('GC=$gc', x['shape'], x['kernel'], 'step_ms',round(x['step_ms'],2),'frac',round(x['frac_of_dmma_peak'],3))
    except Exception: print('GC=$gc', l[:150].strip())"
done
