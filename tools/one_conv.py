"""Run one decoder-shaped conv (B=16, L = 4.8M / C rows) on the tcgen05 kernel: python tools/one_conv.py C k n_adds n_outs [reps]
(for ncu captures and quick timings of a single shape)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from visinger_b200 import _lib
C, k, na, no = (int(a) for a in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
L = _lib.lib()
L.vsg_debug_set_plan.argtypes = [ctypes.c_int32] * 5
L.vsg_debug_last_ms.restype = ctypes.c_float
dev = torch.device("cuda:0")
B, Lq = 16, 4800000 // C // 16 * 16 // 16
Lq = {256: 5000, 128: 25000, 64: 75000, 32: 150000, 16: 300000}[C]
x = torch.randn(B, Lq, C, device=dev).to(torch.bfloat16)
w = torch.randn(C, C, k) / (C * k) ** 0.5
b = torch.zeros(C)
add = torch.randn(B, Lq, C, device=dev).to(torch.bfloat16)
L.vsg_debug_set_plan(0, 0, -1, -1, reps)
raw = torch.empty_like(add) if no > 1 else None
act = torch.empty_like(add)
rc = L.vsg_debug_conv1d_bf16(x.data_ptr(), w.data_ptr(), b.data_ptr(), add.data_ptr() if na > 0 else None,
                             add.data_ptr() if na > 1 else None, 1.0, None, raw.data_ptr() if raw is not None else None,
                             act.data_ptr(), B, Lq, C, C, k, 1, 3 | (8 if na > 0 else 0), 0)
assert rc == 0, L.vsg_last_error()
print(f"C={C} k={k} adds={na} outs={no}: {float(L.vsg_debug_last_ms()) * 1e3:.1f} us")
