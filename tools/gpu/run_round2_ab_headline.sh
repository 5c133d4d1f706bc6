# Round-2 A/B of the headline-step candidates written at the end of round 1 (all OFF by default): stage sub-batching,
# activation-box multicast across 2-CTA clusters, RVQ v2.  ~6 min of box time.
mkdir -p gpurun_out
# stage-level sub-batching so that a stage's ResBlocks work out of the L2 (host-side only, bit-identical): parity, then A/B
HILCODEC_STAGE_CHUNK_MB=64 timeout 300 python -m pytest tests/test_gpu_codec.py -x -q --tb=short -p no:cacheprovider > gpurun_out/ab_pytest_chunk.log 2>&1
echo "pytest [HILCODEC_STAGE_CHUNK_MB=64] rc=$?"; tail -3 gpurun_out/ab_pytest_chunk.log | cut -c1-300
for mb in 0 32 64 96; do
  HILCODEC_STAGE_CHUNK_MB=$mb timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_bench_chunk_$mb.json 2> gpurun_out/ab_bench_chunk_$mb.err
  echo "bench [HILCODEC_STAGE_CHUNK_MB=$mb] rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/ab_bench_chunk_$mb.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['gpu_launches'], d.get('kernel_categories'))" | cut -c1-600
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
