# round 2, call 24: short-chunk depthwise rows many per CTA, STFT tensor-core tiles from 32 windows, flat tiles for pitch-4 rows
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2c24_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|error|^E  |FAILED" gpurun_out/r2c24_pytest.log | cut -c1-600 | tail -12
for wl in stream64 stream1; do
timeout 600 python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2c24_bench_$wl.json 2> gpurun_out/r2c24_bench_$wl.err
echo "bench $wl rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2c24_bench_$wl.json').read().strip().splitlines()[-1]); print(d['ms_per_hop'], d['value'], d['gpu_launches_per_hop'], 'e2e', d['e2e']['value'])"
done
timeout 300 python bench.py --workload music256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c24_bench_music256.json 2> gpurun_out/r2c24_bench_music256.err
python -c "
import json; d=json.loads(open('gpurun_out/r2c24_bench_music256.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['clocks']['sm_mhz'], d['gpu_launches'], {k: round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()})"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c24_launches_stream64.csv python tools/gpu/stream_hops.py 64 4 > /dev/null 2>&1
echo "ncu rc=$?"
