mkdir -p gpurun_out
timeout 300 python tools/gpu/h_check.py > gpurun_out/h_check.log 2>&1; echo "h_check rc=$?"; tail -30 gpurun_out/h_check.log
for g in f16 tf32; do
  echo "== HILCODEC_GEMM=$g"; HILCODEC_GEMM=$g timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_$g.log; python -c "
import sys,json; d=json.loads(open('gpurun_out/bench_$g.log').read()); print(round(d['ms_per_step'],2), 'gemm', round(d['kernel_categories']['pointwise_gemm']['ms_per_step'],2), 'checksum', d['e2e']['checksum'])"
done
