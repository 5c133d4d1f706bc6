# round 2, call 6: STFT gather span staging, decoder input 1x1 on the tensor pipe, tolerance model -- parity + timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2c6_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|error|^E  |FAILED" gpurun_out/r2c6_pytest.log | cut -c1-600 | tail -12
for i in 1 2; do
timeout 300 python bench.py --workload music256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c6_bench_$i.json 2> gpurun_out/r2c6_bench_$i.err
echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2c6_bench_$i.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['clocks']['sm_mhz'], d['gpu_launches'], {k: round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()})"
done
for wl in speech64 stream1 stream64; do
  timeout 300 python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2c6_$wl.json 2> gpurun_out/r2c6_$wl.err
  echo "$wl rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2c6_$wl.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d.get('ms_per_hop'), d['e2e'].get('ms_per_hop'), d['gpu_launches'])"
done
