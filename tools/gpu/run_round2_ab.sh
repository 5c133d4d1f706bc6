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
# large-batch RVQ v2 (8 frames per warp, one balanced wave): parity, then the headline step with and without it
HILCODEC_RVQ_V2=1 timeout 300 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/ab_pytest_rvq_v2.log 2>&1
echo "pytest [HILCODEC_RVQ_V2=1] rc=$?"; tail -3 gpurun_out/ab_pytest_rvq_v2.log | cut -c1-300
for v in 0 1; do
  HILCODEC_RVQ_V2=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_bench_rvq_v2_$v.json 2> gpurun_out/ab_bench_rvq_v2_$v.err
  echo "bench [HILCODEC_RVQ_V2=$v] rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/ab_bench_rvq_v2_$v.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d.get('kernel_categories'))" | cut -c1-600
done
# activation-box multicast across 2-CTA clusters in the wide layers (gemm_h.cu, HILCODEC_CLUSTER_X=1): parity first (a
# protocol error traps -> launch failure, not a hang; the outer timeout bounds it anyway), then the headline step
HILCODEC_CLUSTER_X=1 timeout 300 python -m pytest tests/test_gpu_codec.py tests/test_gpu_ops.py -x -q --tb=short -p no:cacheprovider > gpurun_out/ab_pytest_cluster_x.log 2>&1
echo "pytest [HILCODEC_CLUSTER_X=1] rc=$?"; tail -3 gpurun_out/ab_pytest_cluster_x.log | cut -c1-300
for v in 0 1; do
  HILCODEC_CLUSTER_X=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_bench_cluster_x_$v.json 2> gpurun_out/ab_bench_cluster_x_$v.err
  echo "bench [HILCODEC_CLUSTER_X=$v] rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/ab_bench_cluster_x_$v.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d.get('kernel_categories'))" | cut -c1-600
done
# config 4 in bench.py's JSON contract (default kernels, then the streaming kernels)
timeout 200 python bench.py --workload stream1 --steps 5 --warmup 3 > gpurun_out/ab_bench_stream1_default.json 2> gpurun_out/ab_bench_stream1_default.err; echo "bench stream1 rc=$?"; cut -c1-400 gpurun_out/ab_bench_stream1_default.json
HILCODEC_SKINNY=1 HILCODEC_RVQ_SPLIT=1 timeout 200 python bench.py --workload stream1 --steps 5 --warmup 3 > gpurun_out/ab_bench_stream1_skinny.json 2> gpurun_out/ab_bench_stream1_skinny.err; echo "bench stream1 skinny rc=$?"; cut -c1-400 gpurun_out/ab_bench_stream1_skinny.json
# stage-level sub-batching so that a stage's ResBlocks work out of the L2 (host-side only, bit-identical): parity, then A/B
HILCODEC_STAGE_CHUNK_MB=64 timeout 300 python -m pytest tests/test_gpu_codec.py -x -q --tb=short -p no:cacheprovider > gpurun_out/ab_pytest_chunk.log 2>&1
echo "pytest [HILCODEC_STAGE_CHUNK_MB=64] rc=$?"; tail -3 gpurun_out/ab_pytest_chunk.log | cut -c1-300
for mb in 0 32 64 96; do
  HILCODEC_STAGE_CHUNK_MB=$mb timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_bench_chunk_$mb.json 2> gpurun_out/ab_bench_chunk_$mb.err
  echo "bench [HILCODEC_STAGE_CHUNK_MB=$mb] rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/ab_bench_chunk_$mb.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['gpu_launches'], d.get('kernel_categories'))" | cut -c1-600
done
