mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest.log | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_music256_tc.log 2>&1; echo "bench256 rc=$?"; tail -1 gpurun_out/bench_music256_tc.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step']); [print(k, round(v['ms_per_step'],2), round(v['tflops'],1), round(v['gbs'])) for k,v in d['kernel_categories'].items()]"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -s 230 -c 115 --csv --log-file gpurun_out/launches_r1_tc.csv python bench.py --steps 1 --warmup 3 --workload speech64 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
