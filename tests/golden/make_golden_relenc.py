"""Golden vectors of the UNMODIFIED reference `RelativeEncoder` (modules/rel_transformer.py:257-320).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_relenc.py

`small_relenc*.npz`: reduced configurations with weights, inputs and reference outputs (per-utterance condition, per-frame
condition, no condition; ragged masks; T not a multiple of the attention tile).  `full_relenc.npz`: the FramePriorNetwork
configuration (192 / 768 / 2 heads / 4 layers / k9, gin 1 per frame; modules/visinger/encoder.py:62-64): seeds and strided
output slices.
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

from modules.rel_transformer import RelativeEncoder  # noqa: E402  (reference)

from oracle import visinger_oracle as O  # noqa: E402

torch.set_grad_enabled(False)


def make_case(name, cfg, seed, B, T, lengths, g_t, store_weights, slice_t=None):
    ref = RelativeEncoder(cfg["hidden"], cfg["filter"], cfg["n_heads"], cfg["n_layers"], kernel_size=cfg["kernel_size"],
                          p_dropout=0.1, window_size=cfg["window"], gin_channels=cfg["gin"] or None).eval()
    sd = O.synth_rel_encoder_state_dict(O.rel_encoder_param_shapes(cfg["hidden"], cfg["filter"], cfg["n_heads"], cfg["n_layers"],
                                                                   cfg["kernel_size"], cfg["window"], cfg["gin"] or None), seed)
    ref.load_state_dict(sd, strict=True)
    gen = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, cfg["hidden"], T, generator=gen)
    g = torch.randn(B, cfg["gin"], T if g_t else 1, generator=gen) if cfg["gin"] else None
    mask = torch.ones(B, 1, T)
    for b, n in enumerate(lengths):
        mask[b, :, n:] = 0
    y = ref(x, mask, g)
    yo = O.rel_encoder(sd, x, mask, g, n_heads=cfg["n_heads"], n_layers=cfg["n_layers"], kernel_size=cfg["kernel_size"],
                       window=cfg["window"])
    y64 = O.rel_encoder({k: v.double() for k, v in sd.items()}, x.double(), mask.double(), None if g is None else g.double(),
                        n_heads=cfg["n_heads"], n_layers=cfg["n_layers"], kernel_size=cfg["kernel_size"], window=cfg["window"])
    err = float((y - yo).abs().max())
    print(f"{name}: oracle vs reference max-abs {err:.3e}; fp32 vs fp64 {float((y.double() - y64).abs().max()):.3e}; "
          f"|y|max {float(y.abs().max()):.3f}")
    assert err <= 2e-5
    out = {"cfg_keys": np.array(sorted(cfg)), "cfg_vals": np.array([cfg[k] for k in sorted(cfg)]), "seed": seed,
           "B": B, "T": T, "lengths": np.array(lengths), "g_t": int(g_t)}
    if store_weights:
        out.update({"w/" + k: v.numpy() for k, v in sd.items()})
        out.update({"x": x.numpy(), "mask": mask.numpy(), "y": y.numpy()})
        if g is not None:
            out["g"] = g.numpy()
    else:
        out.update({"y": y[:, ::7, ::slice_t].numpy(), "slice_t": slice_t})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


if __name__ == "__main__":
    small = dict(hidden=32, filter=64, n_heads=2, n_layers=2, kernel_size=3, window=4, gin=8)
    make_case("small_relenc", small, 21, 3, 150, [150, 97, 5], False, True)
    make_case("small_relenc_gframe", dict(small, gin=1, kernel_size=9, n_layers=3), 22, 2, 70, [70, 33], True, True)
    make_case("small_relenc_nog", dict(small, gin=0, hidden=96, n_heads=2, filter=48), 23, 2, 64, [64, 64], False, True)
    make_case("full_relenc", dict(hidden=192, filter=768, n_heads=2, n_layers=4, kernel_size=9, window=4, gin=1), 24, 2, 200,
              [200, 131], True, False, slice_t=9)
