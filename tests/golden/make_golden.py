"""Generate golden vectors from the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference, which does not exist on the
GPU box):

    python tests/golden/make_golden.py

It imports `ResidualCouplingBlock` (modules/visinger/flow.py:15) and `Generator`
(modules/visinger/decoder.py:13) from /root/reference, loads deterministic synthetic
weights into them through their own `load_state_dict`, runs them on CPU in fp32 and
fp64 and stores

  * `small_*.npz`  -- complete cases on reduced configurations: weights, inputs and
    reference outputs (a few hundred KB each), so the fixture is self-contained;
  * `full_*.npz`   -- the full `config/models/visinger.yaml` configuration: only the
    weight SEED, the inputs' seed and strided slices of the reference outputs (weights
    are rebuilt from the seed by `oracle.visinger_oracle.synth_state_dict`).

The reference ships no tests or golden vectors of its own, so
these reference-generated fixtures are what pins the oracle.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("VISINGER_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

from modules.visinger.flow import ResidualCouplingBlock  # noqa: E402  (reference)
from modules.visinger.decoder import Generator  # noqa: E402  (reference)

from oracle import visinger_oracle as O  # noqa: E402

torch.set_grad_enabled(False)


def _inputs(seed, B, C, T, gin, lengths=None):
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, T, generator=gen)
    g = 0.1 * torch.randn(B, gin, 1, generator=gen) if gin else None
    mask = torch.ones(B, 1, T)
    if lengths is not None:
        for b, n in enumerate(lengths):
            mask[b, :, n:] = 0
    return x, mask, g


def _np(sd):
    return {k: v.numpy() for k, v in sd.items()}


def make_flow_case(name, cfg, seed, B, T, lengths=None, store_weights=True, slice_t=None):
    ref = ResidualCouplingBlock(cfg["channels"], cfg["hidden"], cfg["kernel_size"], cfg["dilation_rate"],
                                cfg["n_layers"], n_flows=cfg["n_flows"], gin_channels=cfg["gin"]).eval()
    shapes = O.flow_param_shapes(cfg["channels"], cfg["hidden"], cfg["kernel_size"], cfg["n_layers"],
                                 cfg["n_flows"], cfg["gin"])
    assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == shapes, "state-dict layout drift"
    sd = O.synth_state_dict(shapes, seed)
    ref.load_state_dict(sd)
    x, mask, g = _inputs(seed + 1, B, cfg["channels"], T, cfg["gin"], lengths)
    x = x * mask
    z_rev = ref(x, mask, g=g, reverse=True)
    z_fwd = ref(x, mask, g=g, reverse=False)
    ref64 = ResidualCouplingBlock(cfg["channels"], cfg["hidden"], cfg["kernel_size"], cfg["dilation_rate"],
                                  cfg["n_layers"], n_flows=cfg["n_flows"], gin_channels=cfg["gin"]).eval().double()
    ref64.load_state_dict({k: v.double() for k, v in sd.items()})
    z_rev64 = ref64(x.double(), mask.double(), g=None if g is None else g.double(), reverse=True)
    kw = dict(channels=cfg["channels"], hidden=cfg["hidden"], kernel_size=cfg["kernel_size"],
              dilation_rate=cfg["dilation_rate"], n_layers=cfg["n_layers"], n_flows=cfg["n_flows"])
    o_rev = O.flow(sd, x, mask, g, reverse=True, **kw)
    o_fwd = O.flow(sd, x, mask, g, reverse=False, **kw)
    print(f"[{name}] oracle vs reference: rev {float((o_rev - z_rev).abs().max()):.3e} "
          f"fwd {float((o_fwd - z_fwd).abs().max()):.3e}; fp32 vs fp64 {float((z_rev - z_rev64).abs().max()):.3e}; "
          f"|z-x| {float((z_rev - x).abs().max()):.3f} |z| {float(z_rev.abs().max()):.3f}")
    out = dict(kind="flow", seed=seed, B=B, T=T, lengths=np.array(lengths if lengths else [T] * B),
               cfg_keys=np.array(list(cfg.keys())), cfg_vals=np.array(list(cfg.values())))
    if store_weights:
        out.update({"w/" + k: v for k, v in _np(sd).items()})
        out.update(x=x.numpy(), mask=mask.numpy(), g=g.numpy() if g is not None else np.zeros(0),
                   z_rev=z_rev.numpy(), z_fwd=z_fwd.numpy(), z_rev64=z_rev64.numpy())
    else:
        sl = slice(None, None, slice_t)
        out.update(slice_t=slice_t, z_rev=z_rev[:, :, sl].numpy(), z_fwd=z_fwd[:, :, sl].numpy(),
                   z_rev64=z_rev64[:, :, sl].numpy())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def make_gen_case(name, cfg, seed, B, T, store_weights=True, slice_t=None):
    args = (cfg["initial_channel"], cfg["resblock"], cfg["rk"], cfg["rd"], cfg["ur"], cfg["uic"], cfg["uk"])
    ref = Generator(*args, gin_channels=cfg["gin"]).eval()
    shapes = O.generator_param_shapes(cfg["initial_channel"], cfg["resblock"], cfg["rk"], cfg["rd"], cfg["ur"],
                                      cfg["uic"], cfg["uk"], cfg["gin"])
    assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == shapes, "state-dict layout drift"
    sd = O.synth_state_dict(shapes, seed)
    ref.load_state_dict(sd)
    x, _, g = _inputs(seed + 1, B, cfg["initial_channel"], T, cfg["gin"])
    wav = ref(x, g=g)
    ref64 = Generator(*args, gin_channels=cfg["gin"]).eval().double()
    ref64.load_state_dict({k: v.double() for k, v in sd.items()})
    wav64 = ref64(x.double(), g=None if g is None else g.double())
    kw = dict(resblock=cfg["resblock"], resblock_kernel_sizes=cfg["rk"], resblock_dilation_sizes=cfg["rd"],
              upsample_rates=cfg["ur"], upsample_kernel_sizes=cfg["uk"])
    o = O.generator(sd, x, g, **kw)
    print(f"[{name}] oracle vs reference: {float((o - wav).abs().max()):.3e}; fp32 vs fp64 "
          f"{float((wav - wav64).abs().max()):.3e}; |wav|max {float(wav.abs().max()):.3f} rms {float(wav.pow(2).mean().sqrt()):.3f}")
    out = dict(kind="generator", seed=seed, B=B, T=T, cfg_json=np.array(repr(cfg)))
    if store_weights:
        out.update({"w/" + k: v for k, v in _np(sd).items()})
        out.update(x=x.numpy(), g=g.numpy() if g is not None else np.zeros(0), wav=wav.numpy(), wav64=wav64.numpy())
    else:
        sl = slice(None, None, slice_t)
        out.update(slice_t=slice_t, wav=wav[:, :, sl].numpy(), wav64=wav64[:, :, sl].numpy())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


FLOW_SMALL = dict(channels=8, hidden=16, kernel_size=5, dilation_rate=1, n_layers=2, n_flows=4, gin=8)
FLOW_SMALL_DIL = dict(channels=4, hidden=8, kernel_size=3, dilation_rate=2, n_layers=3, n_flows=3, gin=0)
FLOW_FULL = dict(channels=192, hidden=192, kernel_size=5, dilation_rate=1, n_layers=4, n_flows=4, gin=256)
GEN_SMALL = dict(initial_channel=8, resblock="1", rk=[3, 7, 11], rd=[[1, 3, 5]] * 3, ur=[5, 3, 2], uic=32,
                 uk=[11, 7, 4], gin=8)
GEN_SMALL_RB2 = dict(initial_channel=8, resblock="2", rk=[3, 5], rd=[[1, 3], [1, 3]], ur=[4, 2], uic=16,
                     uk=[8, 4], gin=0)
GEN_FULL = dict(initial_channel=192, resblock="1", rk=[3, 7, 11], rd=[[1, 3, 5]] * 3, ur=[5, 5, 3, 2, 2], uic=512,
                uk=[11, 11, 7, 4, 4], gin=256)

if __name__ == "__main__":
    make_flow_case("small_flow", FLOW_SMALL, 11, B=3, T=50, lengths=[50, 37, 8])
    make_flow_case("small_flow_dil_odd", FLOW_SMALL_DIL, 12, B=2, T=33, lengths=[33, 20])
    make_gen_case("small_gen", GEN_SMALL, 21, B=2, T=23)
    make_gen_case("small_gen_rb2", GEN_SMALL_RB2, 22, B=2, T=17)
    make_flow_case("full_flow", FLOW_FULL, 1234, B=2, T=300, lengths=[300, 211], store_weights=False, slice_t=7)
    make_gen_case("full_gen", GEN_FULL, 1234, B=1, T=64, store_weights=False, slice_t=13)
