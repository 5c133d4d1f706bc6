# round 2, call 2: parity of the cleaned-up default tree (skinny GEMM / per-stage RVQ / 8-frames-per-warp RVQ promoted,
# never-run variants deleted, clip-pair ResBlocks, range guard, generic RVQ), then the new bench contract and ncu lists
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/r2c2_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|error|config 2:|config 3:|edge cases:|range guard:" gpurun_out/r2c2_pytest.log | cut -c1-400 | tail -20
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2c2_smoke.log
timeout 900 python bench.py > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2c2_bench.json').read().strip().splitlines()[-1])
    print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d.get('cpu_baseline', {}).get('value'), d.get('cpu_baseline', {}).get('kind'))
    print({k: round(v['ms_per_step'], 2) for k, v in d['kernel_categories'].items()})
    print({k: (round(v['frac'], 3), v['bound']) for k, v in d['roofline_by_class'].items()})
    for k, v in d.get('other_workloads', {}).items():
        print(k, {kk: v.get(kk) for kk in ('value', 'ms_per_step', 'ms_per_hop', 'gpu_launches_per_hop', 'error')}, 'e2e', v.get('e2e', {}).get('value'), 'cpu', v.get('cpu_baseline', {}).get('value'), v.get('cpu_baseline', {}).get('one_thread', {}).get('value'))
except Exception as e:
    print('bench parse failed', e)
PY
tail -3 gpurun_out/r2c2_bench.err
HILCODEC_RB_PAIR=0 timeout 300 python bench.py --workload music256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c2_bench_nopair.json 2> gpurun_out/r2c2_bench_nopair.err
echo "bench nopair rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2c2_bench_nopair.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], {k: round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()})"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c2_bench_ref.json 2> gpurun_out/r2c2_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2c2_bench_ref.json
# ncu launch lists: one music256 step (+ the library's own launch manifest of the same step), three streaming hops
HILCODEC_DUMP_LAUNCHES=gpurun_out/r2c2_manifest_music256.json timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.avg.per_cycle_elapsed --clock-control none --csv --log-file gpurun_out/r2c2_launches_music256.csv python bench.py --workload music256 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2c2_ncu_bench.log 2>&1
echo "ncu music256 rc=$?"; grep -c gpu__time_duration gpurun_out/r2c2_launches_music256.csv
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2c2_launches_stream1.csv python tools/gpu/stream_hops.py 1 6 > gpurun_out/r2c2_ncu_stream.log 2>&1
echo "ncu stream1 rc=$?"; grep -c gpu__time_duration gpurun_out/r2c2_launches_stream1.csv
