"""Structured probe of the tcgen05 GEMM operand layouts."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from hilcodec_b200 import _lib
lib = _lib.load()
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = torch.cuda.current_stream().cuda_stream
torch.set_printoptions(linewidth=200, sci_mode=False)

def gemm(w, x, tc=1):
    M, K = w.shape[:2]; B, _, T = x.shape
    xd = x.cuda(); y = torch.full((B, M, T), -7.0, device="cuda")
    lib.hil_set_tensor_cores(tc)
    _lib.check(lib.hil_op_pointwise(P(xd), P(w.contiguous()), None, None, P(y), B, M, K, T, 0, 1.0, st))
    torch.cuda.synchronize()
    return y.cpu()

for K in (32, 64):
    M, T = 128, 128
    x = (torch.arange(K).view(1, K, 1) * 128 + torch.arange(T).view(1, 1, T)).float()
    w = torch.zeros(M, K, 1)
    for m in range(M):
        w[m, m % K, 0] = 1.0
    y = gemm(w, x)[0]
    ref = gemm(w, x, 0)[0]
    print(f"K={K}: ffma ok {torch.equal(ref, x[0][torch.arange(M) % K])}; tc equal {torch.equal(y, ref)}")
    kk = (y // 128).long(); tt = (y % 128).long()
    for m in (0, 1, 2, 7, 8, 9, 31, 32, 33, 64, 127):
        print(f" m={m:3d} expect k={m%K:2d}: got k(t=0..7)={kk[m,:8].tolist()} t'={tt[m,:8].tolist()}  | t=32..35 k={kk[m,32:36].tolist()} t'={tt[m,32:36].tolist()}")
    print(" raw y[0,:8]", y[0, :8].tolist(), " y[1,:8]", y[1, :8].tolist())
# all-ones weights: Y[m][t] = sum_k X[k][t]
K, M, T = 32, 128, 128
x = torch.zeros(1, K, T); x[0, 3, :] = torch.arange(T).float()
w = torch.ones(M, K, 1)
y = gemm(w, x)[0]
print("ones-W, X row3 = t:", y[0, :12].tolist(), y[5, 30:36].tolist())
x = torch.zeros(1, K, T); x[0, :, 5] = torch.arange(K).float() + 1
w = torch.zeros(M, K, 1); w[:, 9, 0] = 1
y = gemm(w, x)[0]
print("W picks k=9, X col5 = k+1: y[0,:8]", y[0, :8].tolist(), "nonzero cols", torch.nonzero(y[0]).flatten().tolist()[:10])
