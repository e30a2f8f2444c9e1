"""A/B of a per-call environment switch of the library (default VSG_CHAIN_MIN_C: resblock chains of a stage on separate
streams only for stages at least that many channels wide) on the bf16 Generator at bench size; each setting is captured
as one CUDA graph, replays interleaved.   python tools/ab_chain_min_c.py [ENV_NAME value ...]   ("unset" = variable absent)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from visinger_b200 import Generator
from visinger_b200.configs import VISINGER_GENERATOR as cfg
vals = sys.argv[2:] or ["0", "128", "256", "1000"]
env_name = sys.argv[1] if len(sys.argv) > 1 else "VSG_CHAIN_MIN_C"
B, T, reps, rounds = 16, 1000, 10, 3
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = Generator(cfg["initial_channel"], cfg["resblock"], cfg["rk"], cfg["rd"], cfg["ur"], cfg["uic"], cfg["uk"],
              gin_channels=cfg["gin"], precision="bf16").to(dev).eval()
x = torch.randn(B, 192, T, device=dev)
g = 0.1 * torch.randn(B, 256, 1, device=dev)
graphs, outs = {}, {}
for v in vals:
    if v == "unset":
        os.environ.pop(env_name, None)
    else:
        os.environ[env_name] = v
    for _ in range(2):
        m(x, g=g)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        outs[v] = m(x, g=g)
    graphs[v] = gr
best = {v: 1e9 for v in vals}
for r in range(rounds):
    for v in vals:
        graphs[v].replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            graphs[v].replay()
        e1.record(); torch.cuda.synchronize()
        best[v] = min(best[v], e0.elapsed_time(e1) / reps)
for v in vals:
    print(f"{env_name}={v}: decoder {best[v]:.3f} ms, identical: {bool(torch.equal(outs[v], outs[vals[0]]))}")
