"""Config 4: frame-by-frame streaming latency, eager launches vs CUDA-graph replay."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from hilcodec_b200 import streaming as S, weights as W
cfg = W.HIL_MUSIC
w = W.load_pretrained("hil_music") if W.have_pretrained("hil_music") else W.random_weights(cfg, 0)
m = S.HILCodec.from_weights(w, 12).cuda()
for B in (1, 64):
    x = (0.1 * torch.randn(B, 1, 320 * 300, device="cuda")).clamp(-1, 1)
    for mode in ("eager", "graph"):
        st = m.new_stream_state(B)
        def run(f):
            chunk = x[:, :, f * 320:(f + 1) * 320]
            return m.codec_forward(chunk, 12, state=st) if mode == "eager" else st.step(chunk, 12)
        for f in range(20): run(f)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for f in range(20, 300): run(f)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 280
        print(f"B={B} {mode}: {dt*1e3:.3f} ms/frame -> {B/dt:.0f} frames/s, {B*(1/75)/dt:.1f}x real time", flush=True)
