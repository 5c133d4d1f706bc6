# Round-2 first call: validate and measure the streaming-latency kernels written (CPU-emulation verified, never run on a
# GPU) at the end of round 1.  ~2 min of box time.
mkdir -p gpurun_out
for cfg in "HILCODEC_SKINNY=1" "HILCODEC_SKINNY=1 HILCODEC_RVQ_SPLIT=1" "HILCODEC_SKINNY=1 HILCODEC_RVQ_SPLIT=1 HILCODEC_SKINNY_PREFER=1"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout 300 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/ab_pytest_$tag.log 2>&1
  echo "pytest [$cfg] rc=$?"; tail -3 gpurun_out/ab_pytest_$tag.log | cut -c1-300
done
for cfg in "HILCODEC_NONE=1" "HILCODEC_SKINNY=1" "HILCODEC_SKINNY=1 HILCODEC_RVQ_SPLIT=1" "HILCODEC_SKINNY=1 HILCODEC_RVQ_SPLIT=1 HILCODEC_SKINNY_PREFER=1" "HILCODEC_SKINNY=1 HILCODEC_RVQ_SPLIT=1 HILCODEC_SKINNY_MAXN=4096"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout 200 python tools/gpu/stream_time.py > gpurun_out/ab_stream_$tag.jsonl 2>&1
  echo "stream [$cfg] rc=$?"; grep -o '"streams": [0-9]*, "mode": "[a-z]*", "frames_timed": [0-9]*, "ms_per_frame": [0-9.]*' gpurun_out/ab_stream_$tag.jsonl
done
# config 4 in bench.py's JSON contract (default kernels, then the streaming kernels)
timeout 200 python bench.py --workload stream1 --steps 5 --warmup 3 > gpurun_out/ab_bench_stream1_default.json 2> gpurun_out/ab_bench_stream1_default.err; echo "bench stream1 rc=$?"; cut -c1-400 gpurun_out/ab_bench_stream1_default.json
HILCODEC_SKINNY=1 HILCODEC_RVQ_SPLIT=1 timeout 200 python bench.py --workload stream1 --steps 5 --warmup 3 > gpurun_out/ab_bench_stream1_skinny.json 2> gpurun_out/ab_bench_stream1_skinny.err; echo "bench stream1 skinny rc=$?"; cut -c1-400 gpurun_out/ab_bench_stream1_skinny.json
