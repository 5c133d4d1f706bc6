mkdir -p gpurun_out
cp hilcodec_b200/libhilcodec_b200.so /tmp/lib_main.so
cp hilcodec_b200/alt/libob4.so hilcodec_b200/libhilcodec_b200.so
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_codec.py -m gpu -q --tb=line -p no:cacheprovider -x > gpurun_out/pytest_ob4.log 2>&1; echo "pytest ob4 rc=$?"; tail -3 gpurun_out/pytest_ob4.log | cut -c1-250
for v in 4 2 4 2; do
  if [ $v -eq 2 ]; then cp /tmp/lib_main.so hilcodec_b200/libhilcodec_b200.so; else cp hilcodec_b200/alt/libob4.so hilcodec_b200/libhilcodec_b200.so; fi
  echo "== OUT_BUFS=$v"; timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_ob$v.json; python -c "
import sys,json; d=json.loads(open('gpurun_out/bench_ob$v.json').read()); print(round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()}, d['e2e']['checksum'])" 2>&1 | tail -1
done
cp /tmp/lib_main.so hilcodec_b200/libhilcodec_b200.so
