"""Run the fused DWS operator at one layer shape (for ncu) and check it against torch."""
import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn.functional as F
from hilcodec_b200 import _lib
lib = _lib.load()
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = torch.cuda.current_stream().cuda_stream
B, Cc, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
skip_on = int(sys.argv[4]); pre = int(sys.argv[5]); post = int(sys.argv[6]); reps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
g = torch.Generator().manual_seed(0)
x = torch.randn(B, Cc, T, generator=g); w = (torch.randn(Cc, Cc, 1, generator=g) / Cc ** 0.5).contiguous()
wd = torch.randn(Cc, 1, 5, generator=g) / 5 ** 0.5; bd = torch.randn(Cc, generator=g); cache = torch.randn(B, Cc, 4, generator=g)
sk = torch.randn(B, Cc, T, generator=g) if skip_on else None
xd, wdd, bdd, cd = x.cuda(), wd.cuda(), bd.cuda(), cache.cuda()
skd = sk.cuda() if skip_on else None
tmp = torch.empty(B, Cc, T, device="cuda"); y = torch.empty(B, Cc, T, device="cuda"); co = torch.empty(B, Cc, 4, device="cuda")
for _ in range(reps):
    _lib.check(lib.hil_op_dws(P(xd), P(w), P(wdd), P(bdd), P(cd), P(co), P(skd), P(tmp), P(y), B, Cc, T, pre, 0.8660254, post, 0.7071, st))
torch.cuda.synchronize()
if B * Cc * T <= 4 * 96 * 24000:
    xp = F.elu(x * (0.8660254 if pre == 2 else 1.0)) if pre else x
    pw = F.conv1d(xp.double(), w.double())
    xin = torch.cat((cache.double(), pw), 2)
    ref = F.conv1d(xin, wd.double(), bd.double(), groups=Cc)
    if skip_on: ref = ref + sk.double()
    if post: ref = F.elu(ref * (0.7071 if post == 2 else 1.0))
    print("max err", (y.cpu().double() - ref).abs().max().item(), "cache err", (co.cpu().double() - xin[:, :, -4:]).abs().max().item())
