"""BASELINE config 4: hil_music frame-by-frame streaming (hop 320, GPU-resident caches), 30 s clip = 2250
sequential frames, B = 1 stream and B = 64 streams; eager launches vs CUDA-graph replay (`StreamState.step`).
Wall-clock around the whole sequential loop (the latency a caller sees), one JSON line per case."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from hilcodec_b200 import streaming as S, weights as W

cfg = W.HIL_MUSIC
pre = W.have_pretrained("hil_music")
w = W.load_pretrained("hil_music") if pre else W.random_weights(cfg, 0)
m = S.HILCodec.from_weights(w, 12).cuda()
FRAMES, WARM = 2250, 20
for B in (1, 64):
    x = (0.1 * torch.randn(B, 1, 320 * (FRAMES + WARM), device="cuda")).clamp(-1, 1)
    for mode in ("eager", "graph"):
        n = FRAMES if mode == "graph" else 300
        st = m.new_stream_state(B)

        def run(f):
            chunk = x[:, :, f * 320:(f + 1) * 320]
            return m.codec_forward(chunk, 12, state=st) if mode == "eager" else st.step(chunk, 12)

        for f in range(WARM):
            run(f)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for f in range(WARM, WARM + n):
            run(f)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        print(json.dumps({"workload": "configs[3]: hil_music streaming, hop 320, per-frame causal cache", "streams": B,
                          "mode": mode, "frames_timed": n, "ms_per_frame": dt * 1e3, "frames_per_s": B / dt,
                          "x_realtime_per_stream": (1 / 75) / dt, "weights": "published" if pre else "random-init"}),
              flush=True)

# Where a frame goes: the library's own event-bracketed launch profile (hil_profile_begin / hil_profile_end, the
# categories bench.py reports) over 50 eager frames of one stream.  Per-launch event pairs add ~2 us each, so the SHARES
# matter, not the sum.
import ctypes as C

from hilcodec_b200 import _lib

lib = _lib.load()
CATS = ["pointwise_gemm_narrow", "stft_gemm", "depthwise", "depthwise_transposed", "conv_pre", "conv_post_tanh", "rvq", "misc",
        "pointwise_gemm_wide", "resblock_fused"]
for B in (1, 64):
    x = (0.1 * torch.randn(B, 1, 320 * 80, device="cuda")).clamp(-1, 1)
    st = m.new_stream_state(B)
    for f in range(20):
        m.codec_forward(x[:, :, f * 320:(f + 1) * 320], 12, state=st)
    torch.cuda.synchronize()
    n_cat = len(CATS)
    ms, fl, by, cnt = (C.c_double * n_cat)(), (C.c_double * n_cat)(), (C.c_double * n_cat)(), (C.c_int64 * n_cat)()
    _lib.check(lib.hil_profile_begin())
    for f in range(20, 70):
        m.codec_forward(x[:, :, f * 320:(f + 1) * 320], 12, state=st)
    _lib.check(lib.hil_profile_end(ms, fl, by, cnt, n_cat))
    print(json.dumps({"streams": B, "profile_frames": 50,
                      "per_frame": {c: {"ms": ms[i] / 50, "launches": cnt[i] // 50} for i, c in enumerate(CATS) if cnt[i]}}),
          flush=True)
