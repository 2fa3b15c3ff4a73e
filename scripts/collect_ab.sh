# A/B of the collect kernel's accumulator staging (1 stage: up to 8 CTAs per SM by tensor memory; 2 stages: 4)
for acc in 1 2; do
  CMH_COLLECT_ACC_STAGES=$acc timeout 300 python bench.py --steps 10 --warmup 3 --no-encode --no-c2 > gpurun_out/r2p_bench_acc$acc.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2p_bench_acc$acc.json').read().splitlines()[-1])
print('acc stages', $acc, 'C4-64 %.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in d['stage_ms'].items()}, d['parity_check']['equal'], [(s['workload'], round(s['ms_per_step'],3), s['parity_check']['equal']) for s in d['sweep']])
PY
done
