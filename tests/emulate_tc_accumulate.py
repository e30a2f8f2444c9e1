"""What would a tensor-core flow at the fp32 tolerance cost in accuracy?  (analysis script, not a test)

tools/acc_probe.py measured on B200 that tcgen05.mma kind::f16 accumulates in fp32 with TRUNCATION, once per
instruction (K = 16): -17 ulp bias on an all-positive K = 960 reduction, rel-L2 9e-7 on random-sign data (CPU fp32: 9e-8).
This script restates that adder on the CPU and runs the oracle's flow (modules/visinger/flow.py:33-40) through it:
every convolution over time is  acc = trunc_fp32(acc + sum_16 x_hi * w_hi)  in the kernel's K order (channel chunk of 64,
tap, 16-wide slice), plus the split-operand correction terms accumulated exactly (optimistic) -- i.e. the best a
3-plane split-bf16 ("24-bit operands") tcgen05 flow could do -- and compares z with the fp64 flow.
    python tests/emulate_tc_accumulate.py [B] [T]
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import visinger_oracle as O   # noqa: E402
from helpers import FLOW_FULL, flow_shapes, make_inputs   # noqa: E402

_real_conv1d = F.conv1d


def trunc32(v: torch.Tensor) -> torch.Tensor:
    """fp64 -> fp32, rounding towards zero."""
    f = v.float()
    over = f.double().abs() > v.abs()
    return torch.where(over, torch.nextafter(f, torch.zeros_like(f)), f)


def tc_conv1d(x, w, b=None, stride=1, padding=0, dilation=1, groups=1, chain=True):
    if x.shape[-1] == 1 or x.dtype != torch.float32:        # speaker-condition GEMV / fp64 reference run: as is
        return _real_conv1d(x, w, b, stride, padding, dilation, groups)
    B, Cin, T = x.shape
    Cout, _, k = w.shape
    xp = F.pad(x, (padding, padding)).double()
    xh = xp.float().bfloat16().double()                      # hi plane of the activations
    wh = w.bfloat16().double()
    exact = _real_conv1d(xp, w.double(), None, 1, 0, dilation)                    # all planes, exact
    hh = _real_conv1d(xh, wh, None, 1, 0, dilation)
    corr = exact - hh                                        # hi*mid, mid*hi, ... accumulated apart (and exactly)
    KC = 64 if Cin % 64 == 0 else 32 if Cin % 32 == 0 else 16
    acc = torch.zeros(B, Cout, T)
    for c0 in range(0, Cin, KC):
        for j in range(k):
            for k0 in range(c0, c0 + KC, 16):
                part = _real_conv1d(xh[:, k0:k0 + 16, j * dilation: j * dilation + T], wh[:, k0:k0 + 16, j:j + 1])
                acc = trunc32(acc.double() + part)
    out = acc.double() + corr
    if b is not None:
        out = out + b.double().view(1, -1, 1)
    return out.float()


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 192
    sd = O.synth_state_dict(flow_shapes(FLOW_FULL), 1234)
    x, mask, g = make_inputs(7, B, 192, T, 256, [T] * (B - 1) + [T - 13])
    x = x * mask
    sd64 = {k_: v.double() for k_, v in sd.items()}
    with torch.no_grad():
        z64 = O.flow(sd64, x.double(), mask.double(), g.double(), reverse=True) * mask.double()
        z32 = O.flow(sd, x, mask, g, reverse=True) * mask
        O.F.conv1d = tc_conv1d
        try:
            ztc = O.flow(sd, x, mask, g, reverse=True) * mask
        finally:
            O.F.conv1d = _real_conv1d
    print(f"B={B} T={T} |z| max {float(z64.abs().max()):.2f}")
    print(f"CPU fp32 flow           vs fp64: max-abs {float((z32.double() - z64).abs().max()):.3e}")
    print(f"emulated tcgen05 flow   vs fp64: max-abs {float((ztc.double() - z64).abs().max()):.3e}   (north_star bound 1e-5)")
    print(f"emulated tcgen05 flow   vs CPU fp32: max-abs {float((ztc - z32).abs().max()):.3e}")


if __name__ == "__main__":
    main()
