# round 2, call 8: residual-VQ search of batches on the tensor cores (GEMM per stage + decision kernel)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -p no:cacheprovider -x -k "rvq" > gpurun_out/r2c8_pytest_rvq.log 2>&1
echo "pytest rvq rc=$?"; grep -E "passed|failed|error|^E  |FAILED" gpurun_out/r2c8_pytest_rvq.log | cut -c1-600 | tail -12
timeout 1500 python -m pytest tests/test_gpu_codec.py tests/test_gpu_parity_full.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2c8_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|error|^E  |FAILED" gpurun_out/r2c8_pytest.log | cut -c1-600 | tail -12
for v in 1 0 1 0; do
HILCODEC_RVQ_TC=$v timeout 300 python bench.py --workload music256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c8_bench_$v.json 2> gpurun_out/r2c8_bench_$v.err
echo "bench rvq_tc=$v rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2c8_bench_$v.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['clocks']['sm_mhz'], d['gpu_launches'], {k: round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()})"
done
HILCODEC_RVQ_TC=1 timeout 300 python bench.py --workload speech64 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('speech64', d['ms_per_step'], d['gpu_launches'], {k: round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()})"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rvq_tc_select|gemm_h_kernel|kmajor|chlast_to_ncw" -s 0 -c 400 --csv --log-file gpurun_out/r2c8_rvq_launches.csv python bench.py --workload music256 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r2c8_rvq_launches.csv')) if len(r) > 10 and 'gpu__time_duration' in r[-3] or (len(r) > 10 and 'gpu__time_duration.sum' in r)]
sel = [r for r in rows if 'rvq_tc_select' in r[4]]
print('select launches', len(sel), [r[-1] for r in sel[:14]])
# the GEMM launch right before each select
ids = {r[0]: r for r in rows}
for r in sel[:13]:
    prev = ids.get(str(int(r[0]) - 1))
    if prev: print(prev[4][:40], prev[-1], '->', r[-1])
PY
