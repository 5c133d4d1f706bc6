mkdir -p gpurun_out
for v in 3 0 2 3 0; do
  echo "== L2PROMO=$v"; HILCODEC_L2PROMO=$v timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_promo$v.json; python -c "
import sys,json; d=json.loads(open('gpurun_out/bench_promo$v.json').read()); print(round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()}, d['e2e']['checksum'])" 2>&1 | tail -1
done
