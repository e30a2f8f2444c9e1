"""How exact is the fp32 accumulation of tcgen05.mma kind::f16?  One flow-shaped conv (192 -> 384, k = 5) on operands that ARE
bf16 values, so every product is exact in fp32 and the only error is the accumulation; compared with an fp64 convolution of
the same operands.  Prints the error in units of the result's ulp, signed (a truncating adder shows up as a bias towards zero).
python tools/acc_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from visinger_b200 import _lib

dev = torch.device("cuda:0")
torch.manual_seed(0)
B, L, Cin, Cout, k = 2, 1024, 192, 384, 5
for name, positive in (("random sign", False), ("all positive", True)):
    x = torch.randn(B, L, Cin)
    w = torch.randn(Cout, Cin, k) / (Cin * k) ** 0.5
    if positive:
        x, w = x.abs(), w.abs()
    x = x.to(torch.bfloat16)
    w = w.to(torch.bfloat16).float()
    out = _lib.debug_conv1d_bf16(x.to(dev), w, None, 1, flags=3).cpu().double()
    ref = F.conv1d(x.double().transpose(1, 2), w.double(), padding=(k - 1) // 2).transpose(1, 2)
    f32 = F.conv1d(x.float().transpose(1, 2), w, padding=(k - 1) // 2).transpose(1, 2).double()
    ulp = torch.finfo(torch.float32).eps * ref.abs().clamp_min(1e-30)
    for nm, o in (("tcgen05", out), ("CPU fp32 conv", f32)):
        e = (o - ref)
        es = e * torch.sign(ref)           # > 0: magnitude too large, < 0: too small
        print(f"{name:13s} {nm:14s}: |ref| mean {ref.abs().mean():.3f}  max-abs err {e.abs().max():.3e}  rel-L2 "
              f"{(e.norm() / ref.norm()):.3e}  mean signed err {float((es / ulp).mean()):+.2f} ulp  rms {float((e / ulp).pow(2).mean().sqrt()):.2f} ulp "
              f"(K = {Cin * k})")
