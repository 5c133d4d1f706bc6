"""A few eager streaming hops of hil_music (for an ncu launch list of config 4): argv = streams hops."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from hilcodec_b200 import streaming as S, weights as W

B, hops = int(sys.argv[1]), int(sys.argv[2])
w = W.load_pretrained("hil_music") if W.have_pretrained("hil_music") else W.random_weights(W.HIL_MUSIC, 0)
m = S.HILCodec.from_weights(w, 12).cuda()
m._core.check_range = False
x = (0.1 * torch.randn(B, 1, 320 * hops, device="cuda")).clamp(-1, 1)
st = m.new_stream_state(B)
for f in range(hops):
    m.codec_forward(x[:, :, f * 320:(f + 1) * 320], 12, state=st)
torch.cuda.synchronize()
print("done", B, hops)
