# round 2, call 25: flat tiles in the STFT kernel
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2c25_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|error|^E  |FAILED" gpurun_out/r2c25_pytest.log | cut -c1-600 | tail -12
for wl in stream64 stream1; do
timeout 600 python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2c25_bench_$wl.json 2> gpurun_out/r2c25_bench_$wl.err
echo "bench $wl rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2c25_bench_$wl.json').read().strip().splitlines()[-1]); print(d['ms_per_hop'], d['value'], d['gpu_launches_per_hop'], 'e2e', d['e2e']['value'])"
done
python -c "
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c25_launches_stream64.csv python tools/gpu/stream_hops.py 64 4 > /dev/null 2>&1
echo "ncu rc=$?"
