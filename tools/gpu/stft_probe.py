import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from hilcodec_b200 import _lib
from hilcodec_b200.weights import dft_basis
lib = _lib.load()
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = torch.cuda.current_stream().cuda_stream
B, n_fft, hop, T = [int(a) for a in sys.argv[1:5]]
L = (T - 1) * hop + n_fft
wav = (0.1 * torch.randn(B, 1, L)).cuda()
w = torch.from_numpy(dft_basis(n_fft)).contiguous()
y = torch.empty(B, n_fft // 2 + 1, T, device="cuda")
for _ in range(3):
    _lib.check(lib.hil_op_stft_logmag(P(wav), P(w), P(y), B, n_fft, hop, T, st))
torch.cuda.synchronize()
