"""Time the fused pair kernel against the two un-fused convs at bench size (B=16)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from visinger_b200 import _lib
L = _lib.lib()
L.vsg_debug_set_plan.argtypes = [ctypes.c_int32] * 5
L.vsg_debug_last_ms.restype = ctypes.c_float
dev = torch.device("cuda:0")
shapes = ((32, 150000), (16, 300000))
if len(sys.argv) > 1:
    shapes = tuple((int(a), 4800000 // int(a)) for a in sys.argv[1:])
for C, Lq in shapes:
    for k, d1 in ((3, 1), (7, 3), (11, 5)):
        B = 16
        x = torch.randn(B, Lq, C, device=dev).to(torch.bfloat16)
        w = torch.randn(C, C, k) / (C * k) ** 0.5
        b = torch.zeros(C)
        L.vsg_debug_set_plan(0, 0, -1, -1, 5)
        out, raw, act = _lib.debug_pair_bf16(x, w, b, w, b, d1, add0=x, want_f32=False, want_raw=False)
        ms_pair = float(L.vsg_debug_last_ms())
        o1 = _lib.debug_conv1d_bf16(x, w, b, d1, flags=3, want_bf16=True, want_f32=False, want_raw=False)
        ms1 = float(L.vsg_debug_last_ms())
        o2 = _lib.debug_conv1d_bf16(x, w, b, 1, flags=3 | 8, add0=x, want_bf16=True, want_f32=False, want_raw=False)
        ms2 = float(L.vsg_debug_last_ms())
        print(f"C={C} k={k} d1={d1}: pair {ms_pair*1e3:.0f} us | unfused c1 {ms1*1e3:.0f} + c2 {ms2*1e3:.0f} = {(ms1+ms2)*1e3:.0f} us", flush=True)
        del x, out, raw, act, o1, o2
        torch.cuda.empty_cache()
