"""Golden vectors of the UNMODIFIED reference `PosteriorEncoder` (modules/visinger/encoder.py:76-101).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_posterior.py

Writes `small_posterior.npz` (reduced configuration: weights, inputs, injected noise, reference outputs) and
`full_posterior.npz` (the model's configuration, models/visinger.py:59-60: seeds and strided output slices).  The
reference draws its noise with torch.randn_like inside forward (encoder.py:97); the script seeds the global generator,
replays the same draw to recover the noise tensor and stores it, so the fixture pins z_q as well as mu_q / logs_q.
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

from modules.visinger.encoder import PosteriorEncoder  # noqa: E402  (reference)

from oracle import visinger_oracle as O  # noqa: E402

torch.set_grad_enabled(False)


def make_case(name, cfg, seed, B, T, lengths, store_weights, slice_t=None):
    ref = PosteriorEncoder(cfg["in_channels"], cfg["out_channels"], cfg["hidden"], cfg["kernel_size"], cfg["dilation_rate"],
                           cfg["n_layers"], cfg["gin"]).eval()
    sd = O.synth_state_dict(O.posterior_param_shapes(cfg["in_channels"], cfg["out_channels"], cfg["hidden"], cfg["kernel_size"],
                                                     cfg["n_layers"], cfg["gin"]), seed)
    missing, unexpected = ref.load_state_dict(sd, strict=True)
    gen = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, cfg["in_channels"], T, generator=gen)
    g = 0.1 * torch.randn(B, cfg["gin"], 1, generator=gen) if cfg["gin"] else None
    mask = torch.ones(B, 1, T)
    for b, n in enumerate(lengths):
        mask[b, :, n:] = 0
    torch.manual_seed(seed + 2)
    z, mu, logs = ref(x, mask, g=g)
    torch.manual_seed(seed + 2)
    noise = torch.randn_like(mu)                     # the draw forward made (the only use of the global generator)
    zo, muo, logso = O.posterior_encoder(sd, x, mask, g, noise, out_channels=cfg["out_channels"], hidden=cfg["hidden"],
                                         kernel_size=cfg["kernel_size"], dilation_rate=cfg["dilation_rate"],
                                         n_layers=cfg["n_layers"])
    err = max(float((z - zo).abs().max()), float((mu - muo).abs().max()), float((logs - logso).abs().max()))
    print(f"{name}: oracle vs reference max-abs {err:.3e}; |z|max {float(z.abs().max()):.3f}")
    assert err == 0.0
    out = {"cfg_keys": np.array(sorted(cfg)), "cfg_vals": np.array([cfg[k] for k in sorted(cfg)]), "seed": seed,
           "B": B, "T": T, "lengths": np.array(lengths)}
    if store_weights:
        out.update({"w/" + k: v.numpy() for k, v in sd.items()})
        out.update({"x": x.numpy(), "mask": mask.numpy(), "noise": noise.numpy(), "z": z.numpy(), "mu": mu.numpy(),
                    "logs": logs.numpy()})
        if g is not None:
            out["g"] = g.numpy()
    else:
        sl = slice(None, None, slice_t)
        out.update({"z": z[:, ::7, sl].numpy(), "mu": mu[:, ::7, sl].numpy(), "logs": logs[:, ::7, sl].numpy(),
                    "slice_t": slice_t})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


if __name__ == "__main__":
    make_case("small_posterior", dict(in_channels=81, out_channels=32, hidden=48, kernel_size=5, dilation_rate=2, n_layers=3,
                                      gin=24), 11, 3, 57, [57, 40, 9], True)
    make_case("full_posterior", dict(in_channels=1025, out_channels=192, hidden=192, kernel_size=5, dilation_rate=1,
                                     n_layers=16, gin=256), 12, 2, 96, [96, 61], False, slice_t=5)
