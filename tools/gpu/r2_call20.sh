# round 2, call 20: where a 64-stream hop goes (launch list of one eager hop + bench categories)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c20_launches_stream64.csv python tools/gpu/stream_hops.py 64 4 > /dev/null 2>&1
echo "ncu rc=$? lines $(wc -l < gpurun_out/r2c20_launches_stream64.csv)"
timeout 600 python bench.py --workload stream64 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2c20_bench_stream64.json 2> gpurun_out/r2c20_bench_stream64.err
python -c "
import json; d=json.loads(open('gpurun_out/r2c20_bench_stream64.json').read().strip().splitlines()[-1]); print(d['ms_per_hop'], d['gpu_launches_per_hop'], {k: round(v['ms_per_step'],3) for k,v in d['kernel_categories'].items()})"
