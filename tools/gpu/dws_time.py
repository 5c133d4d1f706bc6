"""Time the DWS operator (fused vs unfused) at the codec's layer shapes."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from hilcodec_b200 import _lib
lib = _lib.load()
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = torch.cuda.current_stream().cuda_stream
def run(B, Cc, T, skip_on, pre, post, reps=10):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, Cc, T, device="cuda"); w = (torch.randn(Cc, Cc, 1) / Cc ** 0.5).contiguous()
    wd = torch.randn(Cc, 1, 5, device="cuda"); bd = torch.randn(Cc, device="cuda"); cache = torch.randn(B, Cc, 4, device="cuda")
    sk = torch.randn(B, Cc, T, device="cuda") if skip_on else None
    tmp = torch.empty(B, Cc, T, device="cuda"); y = sk if skip_on else torch.empty(B, Cc, T, device="cuda"); co = torch.empty(B, Cc, 4, device="cuda")
    f = lambda: _lib.check(lib.hil_op_dws(P(x), P(w), P(wd), P(bd), P(cache), P(co), P(sk), P(tmp), P(y), B, Cc, T, pre, 0.866, post, 0.7071, st))
    f(); f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
shapes = [(64, 96, 24000), (64, 192, 12000), (64, 64, 24000), (64, 384, 3000), (64, 768, 600)]
for B, Cc, T in shapes:
    gb = B * Cc * T * 4 / 1e6
    for (skip_on, pre, post) in [(0, 0, 0), (1, 0, 0), (0, 1, 1), (1, 0, 2)]:
        us = run(B, Cc, T, skip_on, pre, post)
        traffic = gb * (2 + skip_on)
        print(f"B{B} C{Cc} T{T} skip{skip_on} pre{pre} post{post}: {us:8.1f} us  ({traffic/us*1e-3*1e3/1e3:5.2f} TB/s of {traffic:.0f} MB min traffic)", flush=True)
