mkdir -p gpurun_out
HILCODEC_DIRECT_STORE=1 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_codec.py -m gpu -q --tb=line -p no:cacheprovider -x > gpurun_out/pytest_direct.log 2>&1; echo "pytest direct rc=$?"; tail -3 gpurun_out/pytest_direct.log | cut -c1-250
for v in 1 0 1 0; do
  echo "== DIRECT_STORE=$v"; if [ $v -eq 1 ]; then export HILCODEC_DIRECT_STORE=1; else unset HILCODEC_DIRECT_STORE; fi
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_direct$v.json; python -c "
import sys,json; d=json.loads(open('gpurun_out/bench_direct$v.json').read()); print(round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()}, d['e2e']['checksum'])" 2>&1 | tail -1
done
