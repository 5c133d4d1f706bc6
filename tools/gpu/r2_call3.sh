# round 2, call 3: re-check of the parity tests that failed in call 2 (conditioning-aware bars), the reworked fused
# downsampling epilogue, the streaming-latency changes (RVQ hand-over rows, small-T conv_post, TC tiles for T = 40),
# compute peaks, and full ncu captures of the three kernel classes
mkdir -p gpurun_out
tools/gpu/peaks > gpurun_out/r2c3_peaks.json 2> gpurun_out/r2c3_peaks.err; echo "peaks rc=$?"; cat gpurun_out/r2c3_peaks.json
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/r2c3_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|error|config 2:|config 3:|edge cases:|range guard:|^E  |FAILED" gpurun_out/r2c3_pytest.log | cut -c1-600 | tail -30
for v in 1 0; do
  HILCODEC_FUSE_DOWNSAMPLE=$v timeout 300 python bench.py --workload music256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c3_bench_down$v.json 2> gpurun_out/r2c3_bench_down$v.err
  echo "bench down=$v rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2c3_bench_down$v.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['clocks']['sm_mhz'], {k: round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()})"
done
HILCODEC_STFT_LOGF=1 timeout 300 python bench.py --workload music256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c3_bench_logf.json 2> gpurun_out/r2c3_bench_logf.err
echo "bench logf rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2c3_bench_logf.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['clocks']['sm_mhz'], {k: round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()})"
for cols in 512 0; do
  for wl in stream1 stream64; do
    HILCODEC_TC_MIN_COLS=$cols timeout 300 python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2c3_${wl}_cols$cols.json 2> gpurun_out/r2c3_${wl}_cols$cols.err
    echo "$wl min_cols=$cols rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2c3_${wl}_cols$cols.json').read().strip().splitlines()[-1]); print(d['ms_per_hop'], d['e2e']['ms_per_hop'], d['gpu_launches_per_hop'], {k: round(v['ms'],3) for k,v in d['kernel_categories_per_hop'].items()})"
  done
done
# full ncu captures: dec u2 fused DWS (narrow), dec u1 fused DWS (wide), ResBlock at dec u3, fused downsample stage 0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_h_kernel -s 155 -c 40 -o gpurun_out/r2c3_prof_gemm_h python bench.py --workload music256 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2c3_ncu_gemm_h.log 2>&1
echo "ncu gemm_h rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resblock_kernel -s 21 -c 7 -o gpurun_out/r2c3_prof_rb python bench.py --workload music256 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2c3_ncu_rb.log 2>&1
echo "ncu rb rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -4
