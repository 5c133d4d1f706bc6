mkdir -p gpurun_out
python hilcodec_b200/build.py > /dev/null 2>&1
timeout 120 python tools/gpu/dws_probe.py 2 96 2400 1 0 0 1; timeout 120 python tools/gpu/dws_probe.py 3 64 1000 0 1 1 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/prof_dws_k96 python tools/gpu/dws_probe.py 64 96 24000 1 0 0 3 > gpurun_out/ncu_dws.log 2>&1; echo "ncu rc=$?"
