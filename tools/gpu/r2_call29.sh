# round 2, call 29: ncu --set full of the RVQ decision kernel and of one STFT launch (evidence for profiles/)
mkdir -p gpurun_out
timeout 400 ncu --set full --import-source on --clock-control none -k regex:rvq_tc_select -s 5 -c 1 -o gpurun_out/r2f_rvq_select python bench.py --workload music256 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:stft_tc_kernel -s 1 -c 1 -o gpurun_out/r2f_stft python bench.py --workload music256 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/r2f_rvq_select.ncu-rep gpurun_out/r2f_stft.ncu-rep
