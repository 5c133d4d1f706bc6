mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/prof_dws_k96_b python tools/gpu/dws_probe.py 64 96 24000 1 0 0 3 > gpurun_out/ncu_dws.log 2>&1; echo "ncu rc=$?"
