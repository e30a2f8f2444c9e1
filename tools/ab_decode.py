"""A/B the bf16 Generator at bench size under option masks of vsg_set_tc_options (halo_mode bit field):
    python tools/ab_decode.py 1 134217729 ...        (1 = shipped configuration)
Each mask is timed as a CUDA graph replayed `reps` times, interleaved over `rounds` rounds (same box, same clocks)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from visinger_b200 import Generator, _lib
from visinger_b200.configs import VISINGER_GENERATOR as cfg

masks = [int(a, 0) for a in sys.argv[1:]] or [1]
B, T, reps, rounds = 16, 1000, 10, 3
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = Generator(cfg["initial_channel"], cfg["resblock"], cfg["rk"], cfg["rd"], cfg["ur"], cfg["uic"], cfg["uk"],
              gin_channels=cfg["gin"], precision="bf16").to(dev).eval()
x = torch.randn(B, 192, T, device=dev)
g = 0.1 * torch.randn(B, 256, 1, device=dev)
graphs, outs = {}, {}
for mk in masks:
    _lib.set_tc_options(mk)
    for _ in range(2):
        m(x, g=g)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        outs[mk] = m(x, g=g)
    graphs[mk] = gr
best = {mk: 1e9 for mk in masks}
for r in range(rounds):
    for mk in masks:
        graphs[mk].replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            graphs[mk].replay()
        e1.record()
        torch.cuda.synchronize()
        best[mk] = min(best[mk], e0.elapsed_time(e1) / reps)
for mk in masks:
    d = float((outs[mk] - outs[masks[0]]).abs().max())
    print(f"mask {mk:#x}: decoder {best[mk]:.3f} ms (best of {rounds} x {reps} graph replays), max-abs vs first mask {d:.2e}")
_lib.set_tc_options(1)
