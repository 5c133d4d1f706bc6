# round 2, call 9: resident weights in gemm_h (K <= 192) A/B + parity; RVQ select kernel v3 timing + one ncu --set full
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_codec.py tests/test_gpu_parity_full.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2c9_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|error|^E  |FAILED" gpurun_out/r2c9_pytest.log | cut -c1-600 | tail -12
for v in 1 0 1 0; do
HILCODEC_A_RESIDENT=$v timeout 300 python bench.py --workload music256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c9_bench_$v.json 2> gpurun_out/r2c9_bench_$v.err
echo "bench a_res=$v rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2c9_bench_$v.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['clocks']['sm_mhz'], d['gpu_launches'], {k: round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()})"
done


