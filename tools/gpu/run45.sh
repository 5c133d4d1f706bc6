mkdir -p gpurun_out
for v in 192 0 192 0; do
  echo "== EPI2_MAXK=$v DW only"; HILCODEC_EPI2_DW_ONLY=1 HILCODEC_EPI2_MAXK=$v timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_epidw$v.json; python -c "
import sys,json; d=json.loads(open('gpurun_out/bench_epidw$v.json').read()); print(round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()}, d['e2e']['checksum'])"
done
