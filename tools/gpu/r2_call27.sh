# round 2, call 27: one-kernel ResBlock for 128 < C <= 256 (64-column tiles, two row blocks), re-measured after the issuer work
mkdir -p gpurun_out
HILCODEC_RB_WIDE=1 timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_codec.py tests/test_gpu_parity_full.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2c27_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|error|^E  |FAILED" gpurun_out/r2c27_pytest.log | cut -c1-600 | tail -12
for v in 1 0 1 0; do
HILCODEC_RB_WIDE=$v timeout 300 python bench.py --workload music256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c27_bench_$v.json 2> gpurun_out/r2c27_bench_$v.err
echo "bench rb_wide=$v rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2c27_bench_$v.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['clocks']['sm_mhz'], d['gpu_launches'], {k: round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()})"
done
