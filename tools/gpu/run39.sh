mkdir -p gpurun_out
for v in 0 20 60 150; do
  echo "== HILCODEC_XFORM_SLEEP=$v"; HILCODEC_XFORM_SLEEP=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_sleep$v.json; python -c "
import sys,json; d=json.loads(open('gpurun_out/bench_sleep$v.json').read()); print(round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()})"
done
