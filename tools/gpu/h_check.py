"""A/B check of the fp16-split tensor-core GEMM (gemm_h.cu, mode 17) against the 3xTF32 kernel (mode 1),
the FFMA kernel (mode 0) and fp64: plain pointwise and fused DWS."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F
from hilcodec_b200 import _lib

lib = _lib.load()
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = torch.cuda.current_stream().cuda_stream

def pre_ref(x, pre):
    if pre == 2: x = x * 0.8660254
    if pre: x = F.elu(x)
    return x

def run_pw(B, M, K, T, pre, bias, res, mode, amp=1.0):
    g = torch.Generator().manual_seed(M * 7 + K * 3 + T)
    x = torch.randn(B, K, T, generator=g) * amp
    w = (torch.randn(M, K, 1, generator=g) / K ** 0.5).contiguous()
    b = torch.randn(M, generator=g) if bias else None
    r = torch.randn(B, M, T, generator=g) if res else None
    xp = pre_ref(x, pre)
    y64 = F.conv1d(xp.double(), w.double(), b.double() if bias else None)
    y32 = F.conv1d(xp, w, b)
    if res:
        y64 = y64 + r.double(); y32 = y32 + r
    xd = x.cuda(); bd = b.cuda() if bias else None; rd = r.cuda() if res else None
    y = torch.zeros(B, M, T, device="cuda")
    lib.hil_set_tensor_cores(mode)
    _lib.check(lib.hil_op_pointwise(P(xd), P(w), P(bd), P(rd), P(y), B, M, K, T, pre, 0.8660254, st))
    torch.cuda.synchronize()
    return (y.cpu().double() - y64).abs().max().item(), (y32.double() - y64).abs().max().item()

def run_dws(B, Cc, T, skip, pre, mode):
    g = torch.Generator().manual_seed(Cc + T)
    x = torch.randn(B, Cc, T, generator=g)
    w = (torch.randn(Cc, Cc, 1, generator=g) / Cc ** 0.5).contiguous()
    wd = torch.randn(Cc, 1, 5, generator=g) / 5 ** 0.5
    bd = torch.randn(Cc, generator=g)
    cache = torch.randn(B, Cc, 4, generator=g)
    sk = torch.randn(B, Cc, T, generator=g) if skip else None
    pw = F.conv1d(pre_ref(x, pre).double(), w.double())
    xin = torch.cat((cache.double(), pw), 2)
    ref = F.conv1d(xin, wd.double(), bd.double(), groups=Cc)
    if skip: ref = ref + sk.double()
    xd, wdd, bdd, cd = x.cuda(), wd.cuda(), bd.cuda(), cache.cuda()
    y = sk.cuda() if skip else torch.empty(B, Cc, T, device="cuda")
    tmp = torch.empty(B, Cc, T, device="cuda"); co = torch.empty(B, Cc, 4, device="cuda")
    lib.hil_set_tensor_cores(mode)
    _lib.check(lib.hil_op_dws(P(xd), P(w), P(wdd), P(bdd), P(cd), P(co), P(y if skip else None), P(tmp), P(y), B, Cc, T,
                              pre, 0.8660254, 0, 1.0, st))
    torch.cuda.synchronize()
    return (y.cpu().double() - ref).abs().max().item(), (co.cpu().double() - xin[:, :, -4:]).abs().max().item()

shapes = [(2, 64, 64, 1024, 0, False, False), (2, 96, 96, 2000, 1, True, True), (1, 192, 192, 900, 2, False, False),
          (2, 128, 64, 256, 0, False, False), (2, 64, 33, 640, 0, True, True), (1, 1024, 513, 76, 0, True, True),
          (1, 768, 768, 600, 2, False, False), (3, 384, 384, 132, 1, False, True), (1, 1536, 128, 76, 0, False, False)]
ok = True
for s in shapes:
    try:
        e_h, e_cpu = run_pw(*s, mode=17)
        e_tc, _ = run_pw(*s, mode=1)
        e_ff, _ = run_pw(*s, mode=0)
    except Exception as ex:
        print("EXC", s, ex, flush=True); ok = False; break
    flag = "OK " if e_h < max(16 * e_cpu, 1e-5) else "BAD"
    ok &= flag == "OK "
    print(f"{flag} pw B,M,K,T,pre,bias,res={s}: err h {e_h:.3e} tf32 {e_tc:.3e} ffma {e_ff:.3e} cpu-fp32 {e_cpu:.3e}", flush=True)
# wide dynamic range of the activations (small and large magnitudes)
for amp in (1e-4, 1e-2, 30.0, 1000.0):
    e_h, e_cpu = run_pw(2, 96, 96, 512, 0, False, False, mode=17, amp=amp)
    e_tc, _ = run_pw(2, 96, 96, 512, 0, False, False, mode=1, amp=amp)
    print(f"amp {amp}: err h {e_h:.3e} tf32 {e_tc:.3e} cpu-fp32 {e_cpu:.3e}", flush=True)
    ok &= e_h < max(16 * e_cpu, 1e-5 * amp)
for s in [(2, 96, 2400, True, 1), (3, 64, 1000, False, 2), (1, 192, 248, False, 0), (2, 384, 132, True, 1)]:
    try:
        e_h, c_h = run_dws(*s, mode=17)
        e_tc, c_tc = run_dws(*s, mode=1)
    except Exception as ex:
        print("EXC", s, ex, flush=True); ok = False; break
    flag = "OK " if e_h < 2e-5 and c_h < 1e-5 else "BAD"
    ok &= flag == "OK "
    print(f"{flag} dws B,C,T,skip,pre={s}: err h {e_h:.3e} (cache {c_h:.3e}) tf32 {e_tc:.3e} (cache {c_tc:.3e})", flush=True)
print("ALL OK" if ok else "FAILURES")
