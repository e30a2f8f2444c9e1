"""Time ResidualCouplingBlock.reverse at bench size (B = 16 x T = 1000) in every precision mode: python tools/time_flow.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from visinger_b200 import ResidualCouplingBlock
from visinger_b200.configs import VISINGER_FLOW as cfg

dev = torch.device("cuda:0")
torch.manual_seed(0)
mods = {}
ref = ResidualCouplingBlock(cfg["channels"], cfg["hidden"], cfg["kernel_size"], cfg["dilation_rate"], cfg["n_layers"],
                            n_flows=cfg["n_flows"], gin_channels=cfg["gin"], precision="fp32")
for f in ref.flows:
    if hasattr(f, "post"):
        torch.nn.init.normal_(f.post.weight, 0, 0.05)
        torch.nn.init.normal_(f.post.bias, 0, 0.05)
sd = ref.state_dict()
B, T = 16, 1000
x = torch.randn(B, 192, T, device=dev)
mask = torch.ones(B, 1, T, device=dev)
g = 0.1 * torch.randn(B, 256, 1, device=dev)
outs = {}
for prec in ("fp32", "bf16x3", "bf16"):
    m = ResidualCouplingBlock(cfg["channels"], cfg["hidden"], cfg["kernel_size"], cfg["dilation_rate"], cfg["n_layers"],
                              n_flows=cfg["n_flows"], gin_channels=cfg["gin"], precision=prec)
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    for _ in range(3):
        z = m(x, mask, g=g, reverse=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        z = m(x, mask, g=g, reverse=True)
    e1.record()
    torch.cuda.synchronize()
    outs[prec] = z
    print(f"flow reverse B16 x T1000 {prec:7s}: {e0.elapsed_time(e1) / 10:.3f} ms   max-abs vs fp32 path "
          f"{float((z - outs['fp32']).abs().max()):.3e}")
