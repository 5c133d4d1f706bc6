"""A/B check of the tensor-core (tcgen05, 3xTF32) pointwise GEMM against the FFMA kernel and fp64."""
import ctypes as C
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F
from hilcodec_b200 import _lib

lib = _lib.load()
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = torch.cuda.current_stream().cuda_stream

def run(B, M, K, T, pre, bias, res, tc):
    g = torch.Generator().manual_seed(M * 7 + K * 3 + T)
    x = torch.randn(B, K, T, generator=g)
    w = (torch.randn(M, K, 1, generator=g) / K ** 0.5).contiguous()
    b = torch.randn(M, generator=g) if bias else None
    r = torch.randn(B, M, T, generator=g) if res else None
    xp = x
    if pre == 2: xp = xp * 0.8660254
    if pre: xp = F.elu(xp)
    y64 = F.conv1d(xp.double(), w.double(), b.double() if bias else None)
    y32 = F.conv1d(xp, w, b)
    if res:
        y64 = y64 + r.double(); y32 = y32 + r
    xd = x.cuda(); bd = b.cuda() if bias else None; rd = r.cuda() if res else None
    y = torch.zeros(B, M, T, device="cuda")
    lib.hil_set_tensor_cores(tc)
    _lib.check(lib.hil_op_pointwise(P(xd), P(w), P(bd), P(rd), P(y), B, M, K, T, pre, 0.8660254, st))
    torch.cuda.synchronize()
    err = (y.cpu().double() - y64).abs().max().item()
    cpu_err = (y32.double() - y64).abs().max().item()
    return err, cpu_err

shapes = [(2, 64, 64, 1024, 0, False, False), (2, 96, 96, 2000, 1, True, True), (1, 192, 192, 900, 2, False, False), (2, 128, 64, 256, 0, False, False), (2, 64, 64, 512, 1, False, False), (1, 96, 96, 1000, 1, False, False),
          (2, 192, 384, 300, 0, True, False), (2, 64, 33, 640, 0, True, True), (1, 1024, 513, 75, 0, True, True),
          (1, 768, 768, 600, 2, False, False), (3, 384, 384, 130, 1, False, True), (1, 1536, 128, 75, 0, False, False)]
ok = True
for s in shapes:
    e_tm, e_cpu = run(*s, tc=9)
    e_tc, _ = run(*s, tc=1)
    e_ff, _ = run(*s, tc=0)
    flag = "OK " if max(e_tc, e_tm) < max(16 * e_cpu, 1e-5) else "BAD"
    ok &= flag == "OK "
    print(f"{flag} B,M,K,T,pre,bias,res={s}: err tm {e_tm:.3e} tc {e_tc:.3e}  ffma {e_ff:.3e}  cpu-fp32 {e_cpu:.3e}", flush=True)
print("ALL OK" if ok else "FAILURES")
