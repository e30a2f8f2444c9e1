"""A/B the whole hot path (vsg_infer: prior sampling -> flow reverse -> decoder) at bench size under option masks of
vsg_set_tc_options (halo_mode bit field; 1 = shipped configuration):
    python tools/ab_hotpath.py [--precision bf16] 1 0x80000001 ...
Each mask is captured as one CUDA graph and replayed, interleaved over the rounds (same box, same clocks)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from visinger_b200 import _lib
from visinger_b200.configs import VISINGER_FLOW, VISINGER_GENERATOR
from visinger_b200.models.visinger import HotPath

args = sys.argv[1:]
prec = "bf16"
if args and args[0] == "--precision":
    prec, args = args[1], args[2:]
masks = [int(a, 0) for a in args] or [1]
B, T, reps, rounds = 16, 1000, 10, 3
dev = torch.device("cuda:0")
hp = HotPath.random_init(VISINGER_FLOW, VISINGER_GENERATOR, dev, precision=prec, seed=0)
gen = torch.Generator(device=dev).manual_seed(1)
mu = torch.randn(B, 192, T, device=dev, generator=gen)
logs = 0.3 * torch.randn(B, 192, T, device=dev, generator=gen) - 1.0
noise = torch.randn(B, 192, T, device=dev, generator=gen)
mask = torch.ones(B, 1, T, device=dev)
g = 0.1 * torch.randn(B, 256, 1, device=dev, generator=gen)
graphs, outs = {}, {}
for mk in masks:
    _lib.set_tc_options(mk)
    for _ in range(2):
        hp.infer(mu, logs, noise, mask, g)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        outs[mk] = hp.infer(mu, logs, noise, mask, g)
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
    dw = float((outs[mk][0] - outs[masks[0]][0]).abs().max())
    dz = float((outs[mk][1] - outs[masks[0]][1]).abs().max())
    print(f"mask {mk:#x}: hot path ({prec}) {best[mk]:.3f} ms (best of {rounds} x {reps} graph replays), vs first mask: "
          f"wav max-abs {dw:.2e}, z max-abs {dz:.2e}")
_lib.set_tc_options(1)
