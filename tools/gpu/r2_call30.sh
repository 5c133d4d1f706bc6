# round 2, call 30: leaner scan loop in the RVQ decision kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_codec.py tests/test_gpu_parity_full.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2c30_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r2c30_pytest.log | cut -c1-300
for i in 1 2; do
timeout 300 python bench.py --workload music256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c30_bench_$i.json 2> gpurun_out/r2c30_bench_$i.err
python -c "
import json; d=json.loads(open('gpurun_out/r2c30_bench_$i.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['clocks']['sm_mhz'], d['gpu_launches'], {k: round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()})"
done
timeout 300 python bench.py --workload speech64 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('speech64', d['ms_per_step'], {k: round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items() if k=='rvq'})"
