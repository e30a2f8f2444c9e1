"""The tensor-core fast path against the fp32 path over odd (B, T) shapes and ragged masks: finite, deterministic, at the
mode's noise floor.   python tools/stress_shapes.py [bf16|bf16x3]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import visinger_oracle as O
from helpers import GEN_FULL, gen_shapes, make_inputs, build_gen, FLOW_FULL, flow_shapes
from visinger_b200 import _lib
from visinger_b200.models.visinger import HotPath
dev = torch.device("cuda:0")
sd = O.synth_state_dict(gen_shapes(GEN_FULL), 5)
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
tol = 2e-2 if prec == "bf16" else 2e-5
m = build_gen(GEN_FULL, sd, dev, precision=prec)
m32 = build_gen(GEN_FULL, sd, dev, precision="fp32")
worst = 0.0
for B, T in [(1, 1), (1, 2), (1, 3), (2, 5), (1, 9), (3, 17), (1, 43), (5, 86), (2, 171), (1, 342), (7, 129), (1, 1707), (16, 257), (3, 1000), (1, 2048), (33, 64)]:
    x, _, g = make_inputs(B * 1000 + T, B, 192, T, 256)
    xd, gd = x.to(dev), g.to(dev)
    fast = m(xd, g=gd)
    ref = m32(xd, g=gd)
    again = m(xd, g=gd)
    rel = float((fast - ref).norm() / ref.norm())
    ok = bool(torch.isfinite(fast).all()) and torch.equal(fast, again)
    worst = max(worst, rel)
    print(f"B={B:3d} T={T:5d}: {prec} vs fp32 rel-L2 {rel:.3e} finite+deterministic={ok}", flush=True)
    assert ok and rel < tol
fsd = O.synth_state_dict(flow_shapes(FLOW_FULL), 5)
hp = HotPath.from_configs(FLOW_FULL, GEN_FULL, fsd, sd, dev, precision=prec)
hp32 = HotPath.from_configs(FLOW_FULL, GEN_FULL, fsd, sd, dev, precision="fp32")
for B, T, lens in [(3, 77, [77, 40, 1]), (2, 513, [513, 300]), (4, 1000, [1000, 999, 512, 3])]:
    mu, mask, g = make_inputs(7 * B + T, B, 192, T, 256, lens)
    gen = torch.Generator().manual_seed(T)
    logs = 0.3 * torch.randn(B, 192, T, generator=gen) - 1.0
    noise = torch.randn(B, 192, T, generator=gen)
    a = [t.to(dev) for t in (mu, logs, noise, mask, g)]
    w, z = hp.infer(*a); w32, z32 = hp32.infer(*a)
    rel = float((w - w32).norm() / w32.norm())
    print(f"hot path B={B} T={T} lens={lens}: wav rel-L2 {rel:.3e}, z max-abs {float((z - z32).abs().max()):.3e}", flush=True)
    assert rel < tol and bool(torch.isfinite(w).all())
print("worst", worst)
