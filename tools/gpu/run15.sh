mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_tc -s 10 -c 1 -o gpurun_out/prof_stft_gather python bench.py --steps 1 --warmup 3 --workload speech64 --no-cpu-baseline > gpurun_out/ncu_stft.log 2>&1; echo "ncu rc=$?"
