mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
cp hilcodec_b200/libhilcodec_b200.so /tmp/lib_xg2.so
for v in 2 1 4 2; do
  if [ $v -eq 2 ]; then cp /tmp/lib_xg2.so hilcodec_b200/libhilcodec_b200.so; else cp hilcodec_b200/alt/libxg$v.so hilcodec_b200/libhilcodec_b200.so; fi
  echo "== XG=$v"; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_xg$v.json; python -c "
import sys,json; d=json.loads(open('gpurun_out/bench_xg$v.json').read()); print(round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()}, d['e2e']['checksum'])"
done
cp /tmp/lib_xg2.so hilcodec_b200/libhilcodec_b200.so
