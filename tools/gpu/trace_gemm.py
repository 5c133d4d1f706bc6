"""One hil_music forward at batch 256 x 24000 with HILCODEC_TRACE=1: gemm_h.cu prints, per launch, the cycles every
warp role spent waiting (per k-block, averaged over the CTAs).  argv: [batch]."""
import os
import sys

os.environ["HILCODEC_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from hilcodec_b200 import streaming as S, weights as W

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
w = W.load_pretrained("hil_music") if W.have_pretrained("hil_music") else W.random_weights(W.HIL_MUSIC, 0)
m = S.HILCodec.from_weights(w, 12).cuda()
m._core.check_range = False
x = (0.1 * torch.randn(B, 1, 24000, device="cuda")).clamp(-1, 1)
for i in range(2):
    sys.stderr.write(f"[trace] ---- forward {i}\n")
    sys.stderr.flush()
    m.codec_forward(x, 12)
    torch.cuda.synchronize()
print("done")
