# round 2, call 31: fused upsampling transform shares the activated boundary input between neighbouring lanes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_codec.py -m gpu -q --tb=short -p no:cacheprovider -x -k "upsample or golden or fixture or oracle_parity or config2 or fused_forward" > gpurun_out/r2c31_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r2c31_pytest.log | cut -c1-300
for i in 1 2; do
timeout 300 python bench.py --workload music256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c31_bench_$i.json 2> gpurun_out/r2c31_bench_$i.err
python -c "
import json; d=json.loads(open('gpurun_out/r2c31_bench_$i.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['clocks']['sm_mhz'], {k: round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()})"
done
