# round 2, call 10: where the gemm_h warp roles wait (cycle counters), RVQ select v4, bench
mkdir -p gpurun_out
timeout 600 python tools/gpu/trace_gemm.py 256 2> gpurun_out/r2c10_trace.log
echo "trace rc=$?"; grep -c "\[trace\]" gpurun_out/r2c10_trace.log
HILCODEC_A_RESIDENT=0 timeout 600 python tools/gpu/trace_gemm.py 256 2> gpurun_out/r2c10_trace_nores.log
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -p no:cacheprovider -x -k "rvq" > gpurun_out/r2c10_pytest_rvq.log 2>&1
echo "pytest rvq rc=$?"; tail -2 gpurun_out/r2c10_pytest_rvq.log
for i in 1 2; do
timeout 300 python bench.py --workload music256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c10_bench_$i.json 2> gpurun_out/r2c10_bench_$i.err
echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2c10_bench_$i.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['clocks']['sm_mhz'], d['gpu_launches'], {k: round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()})"
done
