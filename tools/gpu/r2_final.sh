# round 2, record run: full GPU suite, smoke, default bench line, reference arm, ncu launch lists (+ manifests), two ncu --set full captures
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2f_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err
echo "bench default rc=$?"; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2f_bench_default.json').read().strip().splitlines()[-1])
print('music256', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], 'clocks', d['clocks'])
print('roofline', d['roofline'])
print('cpu', d.get('cpu_baseline'))
print('ref gpu', d.get('reference_on_same_gpu'))
for k, v in d.get('other_workloads', {}).items():
    print(k, {kk: v.get(kk) for kk in ('value', 'ms_per_step', 'ms_per_hop', 'gpu_launches', 'gpu_launches_per_hop')}, 'e2e', (v.get('e2e') or {}).get('value'), 'cpu', (v.get('cpu_baseline') or {}).get('value'), v.get('error'))
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err
echo "bench reference rc=$?"; cut -c1-400 gpurun_out/r2f_bench_reference.json
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
for wl in music256 speech64; do
HILCODEC_DUMP_LAUNCHES=gpurun_out/r2f_manifest_$wl.json timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2f_launches_$wl.csv python bench.py --workload $wl --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/r2f_ncu_$wl.err
echo "ncu list $wl rc=$? lines $(wc -l < gpurun_out/r2f_launches_$wl.csv)"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches_stream1.csv python tools/gpu/stream_hops.py 1 6 > /dev/null 2>&1
echo "ncu list stream1 rc=$? lines $(wc -l < gpurun_out/r2f_launches_stream1.csv)"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_h_kernel -s 30 -c 2 -o gpurun_out/r2f_gemm_h_wide python bench.py --workload music256 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:resblock_kernel -s 2 -c 1 -o gpurun_out/r2f_resblock python bench.py --workload music256 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
