# round 2, call 4: parity suite with the batch-tail criterion, default bench (all workloads), full ncu captures of one
# launch per kernel class (small reports: gpurun_out must stay under 64 MiB)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/r2c4_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|error|config 2:|config 3:|edge cases:|range guard:|^E  |FAILED" gpurun_out/r2c4_pytest.log | cut -c1-700 | tail -30
timeout 900 python bench.py > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2c4_bench.json').read().strip().splitlines()[-1])
    print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d.get('cpu_baseline', {}).get('value'), d.get('reference_on_same_gpu'))
    print({k: round(v['ms_per_step'], 2) for k, v in d['kernel_categories'].items()})
    for k, v in d.get('other_workloads', {}).items():
        print(k, {kk: v.get(kk) for kk in ('value', 'ms_per_step', 'ms_per_hop', 'gpu_launches_per_hop', 'error')}, 'e2e', v.get('e2e', {}).get('value'), 'cpu', v.get('cpu_baseline', {}).get('value'), v.get('cpu_baseline', {}).get('one_thread', {}).get('value'), v.get('reference_on_same_gpu', {}).get('value'))
except Exception as e:
    print('bench parse failed', e)
PY
tail -3 gpurun_out/r2c4_bench.err
NCU="ncu --set full --clock-control none --import-source on"
B="python bench.py --workload music256 --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 $NCU -k regex:gemm_h_kernel -s 153 -c 2 -o gpurun_out/r2c4_prof_dws_u2 $B > gpurun_out/r2c4_ncu1.log 2>&1; echo "ncu dws_u2 rc=$?"
timeout 300 $NCU -k regex:gemm_h_kernel -s 146 -c 2 -o gpurun_out/r2c4_prof_dws_u1 $B > gpurun_out/r2c4_ncu2.log 2>&1; echo "ncu dws_u1 rc=$?"
timeout 300 $NCU -k regex:gemm_h_kernel -s 121 -c 1 -o gpurun_out/r2c4_prof_down_s0 $B > gpurun_out/r2c4_ncu3.log 2>&1; echo "ncu down_s0 rc=$?"
timeout 300 $NCU -k regex:resblock_kernel -s 21 -c 1 -o gpurun_out/r2c4_prof_rb_s0 $B > gpurun_out/r2c4_ncu4.log 2>&1; echo "ncu rb_s0 rc=$?"
timeout 300 $NCU -k regex:resblock_kernel -s 25 -c 1 -o gpurun_out/r2c4_prof_rb_u3 $B > gpurun_out/r2c4_ncu5.log 2>&1; echo "ncu rb_u3 rc=$?"
timeout 300 $NCU -k regex:stft_tc_kernel -s 15 -c 2 -o gpurun_out/r2c4_prof_stft $B > gpurun_out/r2c4_ncu6.log 2>&1; echo "ncu stft rc=$?"
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
