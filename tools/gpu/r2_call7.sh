# round 2, call 7: STFT on the fp16-split path (A/B against 3xTF32), parity
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_codec.py tests/test_gpu_parity_full.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2c7_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|error|^E  |FAILED" gpurun_out/r2c7_pytest.log | cut -c1-600 | tail -12
for v in f16 tf32 f16 tf32; do
HILCODEC_STFT=$v timeout 300 python bench.py --workload music256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c7_bench_$v.json 2> gpurun_out/r2c7_bench_$v.err
echo "bench stft=$v rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2c7_bench_$v.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['clocks']['sm_mhz'], d['gpu_launches'], {k: round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()})"
done
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:stft_tc_kernel -s 15 -c 5 --csv --log-file gpurun_out/r2c7_stft_launches.csv python bench.py --workload music256 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
grep -E "gpu__time_duration|tensor" gpurun_out/r2c7_stft_launches.csv | cut -d, -f5,13-16 | tr -d '"' | head -12
