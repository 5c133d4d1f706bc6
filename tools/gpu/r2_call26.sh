# round 2, call 26: cluster RVQ with 1 / 2 / 4 frames per warp (64 frames: 8 clusters instead of 2)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2c26_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|error|^E  |FAILED" gpurun_out/r2c26_pytest.log | cut -c1-600 | tail -12
for fpw in 0 4 0 4; do
for wl in stream64 stream1; do
HILCODEC_RVQ_FPW=$fpw timeout 600 python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2c26_bench_${wl}_$fpw.json 2> gpurun_out/r2c26_bench_${wl}_$fpw.err
echo "bench $wl fpw=$fpw rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2c26_bench_${wl}_$fpw.json').read().strip().splitlines()[-1]); print(d['ms_per_hop'], d['value'], d['gpu_launches_per_hop'], 'e2e', d['e2e']['value'])"
done
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c26_launches_stream64.csv python tools/gpu/stream_hops.py 64 4 > /dev/null 2>&1
echo "ncu rc=$?"
