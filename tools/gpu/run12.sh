echo "== fused"; timeout 300 python tools/gpu/dws_time.py
echo "== unfused (gemm_tc + dwconv5)"; HILCODEC_DISABLE_DWS_FUSION=1 timeout 300 python tools/gpu/dws_time.py
