mkdir -p gpurun_out
timeout 300 python tools/gpu/tc_check.py > gpurun_out/tc_check.log 2>&1; echo "tc_check rc=$?"; tail -11 gpurun_out/tc_check.log
timeout 1500 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest.log | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_music256_tc.log 2>&1; echo "bench256 rc=$?"; tail -1 gpurun_out/bench_music256_tc.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step']); [print(k, round(v['ms_per_step'],2), round(v['tflops'],1), round(v['gbs'])) for k,v in d['kernel_categories'].items()]"
