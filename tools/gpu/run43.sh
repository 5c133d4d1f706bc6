mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
for v in 256 0 512 256 0; do
  echo "== EPI2_MAXK=$v"; HILCODEC_EPI2_MAXK=$v timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_epi$v.json; python -c "
import sys,json; d=json.loads(open('gpurun_out/bench_epi$v.json').read()); print(round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()}, d['e2e']['checksum'])"
done
