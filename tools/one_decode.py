"""Run ONE bf16 Generator forward at bench size (for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import visinger_oracle as O
from helpers import GEN_FULL, gen_shapes, make_inputs, build_gen
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
dev = torch.device("cuda:0")
m = build_gen(GEN_FULL, O.synth_state_dict(gen_shapes(GEN_FULL), 1234), dev, precision=prec)
x, _, g = make_inputs(0, B, 192, T, 256)
w = m(x.to(dev), g=g.to(dev))
torch.cuda.synchronize()
print("ok", tuple(w.shape), float(w.abs().max()))
