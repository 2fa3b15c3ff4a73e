for b in 4 6 8 12 16; do
  CMH_TC_BLOCKS_PER_SM=$b timeout 300 python bench.py --steps 10 --warmup 3 --no-encode --no-sweep > gpurun_out/r2k_bench_b$b.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2k_bench_b$b.json').read().splitlines()[-1])
c=d['c2_map']
print('blocks/SM', $b, 'C4-64 %.3f'%d['ms_per_step'], {k:round(v,3) for k,v in d['stage_ms'].items()}, '| C2 %.3f'%c['ms_per_step'], {k:round(v,3) for k,v in c['stage_ms'].items()})
PY
done
