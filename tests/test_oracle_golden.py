"""CPU: the oracle (oracle/visinger_oracle.py) against the reference-generated golden vectors.

The fixtures under tests/golden/ were produced by tests/golden/make_golden.py from the reference's own
ResidualCouplingBlock / Generator (imported unmodified from /root/reference).  fp32 results are expected
bit-exact on the same torch build; 2e-6 absorbs a different CPU / MKL-DNN dispatch on another host.
"""
import pytest
import torch

from oracle import visinger_oracle as O
from helpers import (load_npz, flow_cfg_of, gen_cfg_of, weights_of, flow_oracle_kw, gen_oracle_kw, flow_shapes,
                     gen_shapes, make_inputs, maxabs, FLOW_FULL, GEN_FULL)

TOL = 2e-6


@pytest.mark.parametrize("name", ["small_flow", "small_flow_dil_odd"])
def test_flow_small_golden(name):
    z = load_npz(name)
    cfg, sd = flow_cfg_of(z), weights_of(z)
    x, mask = torch.from_numpy(z["x"]), torch.from_numpy(z["mask"])
    g = torch.from_numpy(z["g"]) if cfg["gin"] else None
    with torch.no_grad():
        rev = O.flow(sd, x, mask, g, reverse=True, **flow_oracle_kw(cfg))
        fwd = O.flow(sd, x, mask, g, reverse=False, **flow_oracle_kw(cfg))
        rev64 = O.flow({k: v.double() for k, v in sd.items()}, x.double(), mask.double(),
                       None if g is None else g.double(), reverse=True, **flow_oracle_kw(cfg))
    assert maxabs(rev, torch.from_numpy(z["z_rev"])) <= TOL
    assert maxabs(fwd, torch.from_numpy(z["z_fwd"])) <= TOL
    assert maxabs(rev64, torch.from_numpy(z["z_rev64"])) <= 1e-12
    # the coupling must be non-trivial or the fixture proves nothing
    assert maxabs(rev, x) > 0.05


@pytest.mark.parametrize("name", ["small_gen", "small_gen_rb2"])
def test_generator_small_golden(name):
    z = load_npz(name)
    cfg, sd = gen_cfg_of(z), weights_of(z)
    x = torch.from_numpy(z["x"])
    g = torch.from_numpy(z["g"]) if cfg["gin"] else None
    with torch.no_grad():
        wav = O.generator(sd, x, g, **gen_oracle_kw(cfg))
    hop = 1
    for u in cfg["ur"]:
        hop *= u
    assert wav.shape == (x.shape[0], 1, x.shape[2] * hop)
    assert maxabs(wav, torch.from_numpy(z["wav"])) <= TOL


def test_flow_full_config_golden():
    z = load_npz("full_flow")
    sd = O.synth_state_dict(flow_shapes(FLOW_FULL), int(z["seed"]))
    x, mask, g = make_inputs(int(z["seed"]) + 1, int(z["B"]), 192, int(z["T"]), 256, z["lengths"].tolist())
    x = x * mask
    st = int(z["slice_t"])
    with torch.no_grad():
        rev = O.flow(sd, x, mask, g, reverse=True)
        fwd = O.flow(sd, x, mask, g, reverse=False)
        back = O.flow(sd, rev, mask, g, reverse=False)
    assert maxabs(rev[:, :, ::st], torch.from_numpy(z["z_rev"])) <= TOL
    assert maxabs(fwd[:, :, ::st], torch.from_numpy(z["z_fwd"])) <= TOL
    assert maxabs(rev[:, :, ::st], torch.from_numpy(z["z_rev64"])) <= 1e-5
    assert maxabs(back, x) <= 1e-5      # invariant: forward(reverse(x)) == x


def test_generator_full_config_golden():
    z = load_npz("full_gen")
    sd = O.synth_state_dict(gen_shapes(GEN_FULL), int(z["seed"]))
    x, _, g = make_inputs(int(z["seed"]) + 1, int(z["B"]), 192, int(z["T"]), 256)
    st = int(z["slice_t"])
    with torch.no_grad():
        wav = O.generator(sd, x, g)
    assert wav.shape == (1, 1, 64 * 300)
    assert maxabs(wav[:, :, ::st], torch.from_numpy(z["wav"])) <= TOL
    assert maxabs(wav[:, :, ::st], torch.from_numpy(z["wav64"])) <= 1e-6


def test_prior_sample_and_hot_path_shapes():
    gen = torch.Generator().manual_seed(5)
    B, C, T = 2, 8, 11
    mu, logs, noise = (torch.randn(B, C, T, generator=gen) for _ in range(3))
    mask = torch.ones(B, 1, T)
    mask[1, :, 7:] = 0
    z = O.prior_sample(mu, logs * 0.1, noise, mask)
    assert torch.equal(z[1, :, 7:], torch.zeros(C, 4))
    assert torch.allclose(z[0], mu[0] + noise[0] * torch.exp(0.1 * logs[0]))
