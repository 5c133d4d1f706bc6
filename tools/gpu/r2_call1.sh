# round 2, call 1: default-tree parity with the published hil_music weights on the box, then both A/B scripts
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r2c1_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2c1_pytest_default.log 2>&1
echo "pytest default rc=$?"; tail -5 gpurun_out/r2c1_pytest_default.log | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c1_bench_default.json 2> gpurun_out/r2c1_bench_default.err
echo "bench default rc=$?"; cut -c1-600 gpurun_out/r2c1_bench_default.json
bash tools/gpu/run_round2_ab_streaming.sh
bash tools/gpu/run_round2_ab_headline.sh
