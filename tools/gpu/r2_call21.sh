# round 2, call 21: flat tensor-core tiles for short chunks of many streams (64 streams x 8 samples)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_codec.py tests/test_gpu_parity_full.py tests/test_gpu_next_rows.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2c21_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|error|^E  |FAILED" gpurun_out/r2c21_pytest.log | cut -c1-600 | tail -12
for v in 128 0 128 0; do
HILCODEC_TC_FLAT=$v timeout 600 python bench.py --workload stream64 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2c21_bench_$v.json 2> gpurun_out/r2c21_bench_$v.err
echo "bench flat=$v rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2c21_bench_$v.json').read().strip().splitlines()[-1]); print(d['ms_per_hop'], d['value'], d['gpu_launches_per_hop'], 'e2e', d['e2e']['value'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c21_launches_stream64.csv python tools/gpu/stream_hops.py 64 4 > /dev/null 2>&1
echo "ncu rc=$?"
