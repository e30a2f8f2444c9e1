"""A/B the L2-resident sub-batching of the bf16 Generator (vsg_set_tc_options l2_tensor_mb / min_tiles; VSG_L2_ONLY_C
restricts it to one stage width):  python tools/ab_l2.py MASK L2_MB [MIN_TILES]   -- one configuration per process."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from visinger_b200 import Generator, _lib
from visinger_b200.configs import VISINGER_GENERATOR as cfg

mask = int(sys.argv[1], 0) if len(sys.argv) > 1 else 1
l2 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
mt = int(sys.argv[3]) if len(sys.argv) > 3 else -1
B, T, reps, rounds = 16, 1000, 10, 3
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = Generator(cfg["initial_channel"], cfg["resblock"], cfg["rk"], cfg["rd"], cfg["ur"], cfg["uic"], cfg["uk"],
              gin_channels=cfg["gin"], precision="bf16").to(dev).eval()
x = torch.randn(B, 192, T, device=dev)
g = 0.1 * torch.randn(B, 256, 1, device=dev)
_lib.set_tc_options(mask, 1, l2, mt)
for _ in range(2):
    m(x, g=g)
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    out = m(x, g=g)
best = 1e9
for r in range(rounds):
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / reps)
print(f"mask {mask:#x} l2_tensor_mb {l2} min_tiles {mt} only_c {os.environ.get('VSG_L2_ONLY_C', '-')}: decoder {best:.3f} ms, checksum {float(out.double().abs().sum()):.6f}")
