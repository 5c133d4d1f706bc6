mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=line -p no:cacheprovider -k "upsample_fused" > gpurun_out/pytest_pl.log 2>&1; rc=$?; echo "planes rc=$rc"; tail -8 gpurun_out/pytest_pl.log | cut -c1-250
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python -m pytest tests/test_gpu_codec.py -m gpu -q --tb=line -p no:cacheprovider -x > gpurun_out/pytest_pl2.log 2>&1; echo "codec rc=$?"; tail -3 gpurun_out/pytest_pl2.log | cut -c1-250
for v in 1 0 1 0; do
  echo "== PLANES=$v"; HILCODEC_PLANES=$v timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_pl$v.json; python -c "
import sys,json; d=json.loads(open('gpurun_out/bench_pl$v.json').read()); print(round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()}, d['e2e']['checksum'])" 2>&1 | tail -1
done
