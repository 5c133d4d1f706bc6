mkdir -p gpurun_out
cp hilcodec_b200/libhilcodec_b200.so /tmp/lib_main.so
cp hilcodec_b200/alt/libe640.so hilcodec_b200/libhilcodec_b200.so
HILCODEC_EPI2_MAXK=2048 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_codec.py -m gpu -q --tb=line -p no:cacheprovider -x > gpurun_out/pytest_e640.log 2>&1; echo "pytest e640 rc=$?"; tail -3 gpurun_out/pytest_e640.log | cut -c1-250
for v in "192 1" "256 1" "256 0" "2048 0" "0 0" "192 1" "0 0"; do
  set -- $v
  echo "== EPI2_MAXK=$1 dw_only=$2"; if [ $2 -eq 1 ]; then export HILCODEC_EPI2_DW_ONLY=1; else unset HILCODEC_EPI2_DW_ONLY; fi
  HILCODEC_EPI2_MAXK=$1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_e640.json; python -c "
import sys,json; d=json.loads(open('gpurun_out/bench_e640.json').read()); print(round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()}, d['e2e']['checksum'])" 2>&1 | tail -1
done
cp /tmp/lib_main.so hilcodec_b200/libhilcodec_b200.so
