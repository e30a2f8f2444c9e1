"""Per-launch time of the flow-shaped convolutions at bench size (B16 x T1000 = 16 000 rows: about one 128-row tile per SM),
back to back in one stream: python tools/time_small_conv.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from visinger_b200 import _lib
L = _lib.lib()
dev = torch.device("cuda:0")
B, T, reps = 16, 1000, 50
for Cin, Cout, k, planes in ((192, 192, 1, 1), (192, 384, 1, 1), (192, 384, 5, 1), (96, 192, 1, 1), (192, 192, 1, 3), (192, 384, 5, 3)):
    x = torch.randn(B, T, planes * Cin, device=dev).to(torch.bfloat16)
    w = torch.randn(Cout, Cin, k) / (Cin * k) ** 0.5
    b = torch.zeros(Cout)
    raw = torch.empty(B, T, planes * Cout, device=dev, dtype=torch.bfloat16)
    L.vsg_debug_set_plan(0, 0, -1, -1, reps)
    flags = 3 | (256 if planes == 3 else 0)
    rc = L.vsg_debug_conv1d_bf16(x.data_ptr(), w.data_ptr(), b.data_ptr(), None, None, 1.0, None, raw.data_ptr(), None,
                                 B, T, Cin, Cout, k, 1, flags, 0)
    assert rc == 0, _lib.last_error() if hasattr(_lib, "last_error") else rc
    ms = float(L.vsg_debug_last_ms())
    flop = 2.0 * B * T * Cin * Cout * k * (6 if planes == 3 else 1)
    print(f"conv {Cin}->{Cout} k{k} planes {planes}: {ms * 1e3:.1f} us per launch ({flop / ms / 1e9:.0f} TFLOP/s of MMA work)")
L.vsg_debug_set_plan(0, 0, -1, -1, 1)
