mkdir -p gpurun_out
timeout 300 python tools/gpu/h_check.py > gpurun_out/h_check.log 2>&1; echo "h_check rc=$?"; tail -22 gpurun_out/h_check.log
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
for v in 0 1; do
  echo "== HILCODEC_ELU_POLY=$v"; HILCODEC_ELU_POLY=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_poly$v.json; python -c "
import sys,json; d=json.loads(open('gpurun_out/bench_poly$v.json').read()); print(round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()}, 'checksum', d['e2e']['checksum'])"
done
