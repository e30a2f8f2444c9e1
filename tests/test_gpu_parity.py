"""GPU parity tests: the CUDA path (through the nn.Module mirror -> ctypes -> C ABI) against
  (1) the reference-generated golden vectors in tests/golden/, and
  (2) the CPU oracle run live on the same seeded inputs.

Tolerances (BASELINE.json north_star): fp32 mode -- flow z max-abs <= 1e-5, waveform max-abs <= 1e-4.
"""
import pytest
import torch

from oracle import visinger_oracle as O
from helpers import (load_npz, flow_cfg_of, gen_cfg_of, weights_of, flow_oracle_kw, gen_oracle_kw, flow_shapes,
                     gen_shapes, make_inputs, build_flow, build_gen, maxabs, FLOW_FULL, GEN_FULL)

pytestmark = pytest.mark.gpu

Z_TOL = 1e-5     # flow-inverse z, fp32 mode
WAV_TOL = 1e-4   # waveform, fp32 mode


# ------------------------------------------------------------------ golden vectors (reference-made)
@pytest.mark.parametrize("name", ["small_flow", "small_flow_dil_odd"])
def test_flow_small_golden(cuda_device, name):
    z = load_npz(name)
    cfg, sd = flow_cfg_of(z), weights_of(z)
    m = build_flow(cfg, sd, cuda_device)
    x, mask = torch.from_numpy(z["x"]).to(cuda_device), torch.from_numpy(z["mask"]).to(cuda_device)
    g = torch.from_numpy(z["g"]).to(cuda_device) if cfg["gin"] else None
    rev = m(x, mask, g=g, reverse=True).cpu()
    fwd = m(x, mask, g=g, reverse=False).cpu()
    assert maxabs(rev, torch.from_numpy(z["z_rev"])) <= Z_TOL
    assert maxabs(fwd, torch.from_numpy(z["z_fwd"])) <= Z_TOL
    assert maxabs(rev, torch.from_numpy(z["z_rev64"])) <= Z_TOL


@pytest.mark.parametrize("name", ["small_gen", "small_gen_rb2"])
def test_generator_small_golden(cuda_device, name):
    z = load_npz(name)
    cfg, sd = gen_cfg_of(z), weights_of(z)
    m = build_gen(cfg, sd, cuda_device)
    x = torch.from_numpy(z["x"]).to(cuda_device)
    g = torch.from_numpy(z["g"]).to(cuda_device) if cfg["gin"] else None
    wav = m(x, g=g).cpu()
    assert wav.shape == z["wav"].shape
    assert maxabs(wav, torch.from_numpy(z["wav"])) <= WAV_TOL
    assert maxabs(wav, torch.from_numpy(z["wav64"])) <= WAV_TOL


def test_flow_full_config_golden(cuda_device):
    z = load_npz("full_flow")
    sd = O.synth_state_dict(flow_shapes(FLOW_FULL), int(z["seed"]))
    m = build_flow(FLOW_FULL, sd, cuda_device)
    x, mask, g = make_inputs(int(z["seed"]) + 1, int(z["B"]), 192, int(z["T"]), 256, z["lengths"].tolist())
    x = x * mask
    xd, md, gd = x.to(cuda_device), mask.to(cuda_device), g.to(cuda_device)
    st = int(z["slice_t"])
    rev = m(xd, md, g=gd, reverse=True)
    fwd = m(xd, md, g=gd, reverse=False)
    assert maxabs(rev.cpu()[:, :, ::st], torch.from_numpy(z["z_rev"])) <= Z_TOL
    assert maxabs(fwd.cpu()[:, :, ::st], torch.from_numpy(z["z_fwd"])) <= Z_TOL
    assert maxabs(rev.cpu()[:, :, ::st], torch.from_numpy(z["z_rev64"])) <= Z_TOL
    # invariant the reference implies but never checks: forward(reverse(x)) == x
    back = m(rev, md, g=gd, reverse=False)
    assert maxabs(back.cpu(), x) <= Z_TOL
    # padded frames stay exactly zero (flow masks them)
    assert float(rev[1, :, 211:].abs().max()) == 0.0


def test_generator_full_config_golden(cuda_device):
    z = load_npz("full_gen")
    sd = O.synth_state_dict(gen_shapes(GEN_FULL), int(z["seed"]))
    m = build_gen(GEN_FULL, sd, cuda_device)
    x, _, g = make_inputs(int(z["seed"]) + 1, int(z["B"]), 192, int(z["T"]), 256)
    st = int(z["slice_t"])
    wav = m(x.to(cuda_device), g=g.to(cuda_device)).cpu()
    assert wav.shape == (1, 1, 64 * 300)
    assert maxabs(wav[:, :, ::st], torch.from_numpy(z["wav"])) <= WAV_TOL
    assert maxabs(wav[:, :, ::st], torch.from_numpy(z["wav64"])) <= WAV_TOL


# ------------------------------------------------------------------ live oracle, seeded inputs
@pytest.mark.parametrize("B,T,lengths", [(1, 1, None), (2, 7, [7, 3]), (3, 257, [257, 256, 1]), (4, 400, [400, 333, 120, 17])])
def test_flow_vs_oracle_ragged(cuda_device, B, T, lengths):
    sd = O.synth_state_dict(flow_shapes(FLOW_FULL), 77)
    m = build_flow(FLOW_FULL, sd, cuda_device)
    x, mask, g = make_inputs(100 + T, B, 192, T, 256, lengths)
    x = x * mask
    with torch.no_grad():
        ref = O.flow(sd, x, mask, g, reverse=True)
    got = m(x.to(cuda_device), mask.to(cuda_device), g=g.to(cuda_device), reverse=True).cpu()
    assert maxabs(got, ref) <= Z_TOL


def test_flow_empty_batch_and_zero_length(cuda_device):
    sd = O.synth_state_dict(flow_shapes(FLOW_FULL), 77)
    m = build_flow(FLOW_FULL, sd, cuda_device)
    y = m(torch.zeros(0, 192, 10, device=cuda_device), torch.zeros(0, 1, 10, device=cuda_device),
          g=torch.zeros(0, 256, 1, device=cuda_device), reverse=True)
    assert y.shape == (0, 192, 10)
    y = m(torch.zeros(2, 192, 0, device=cuda_device), torch.zeros(2, 1, 0, device=cuda_device),
          g=torch.zeros(2, 256, 1, device=cuda_device), reverse=True)
    assert y.shape == (2, 192, 0)


@pytest.mark.parametrize("B,T", [(1, 1), (2, 5), (1, 130), (3, 33)])
def test_generator_vs_oracle(cuda_device, B, T):
    sd = O.synth_state_dict(gen_shapes(GEN_FULL), 78)
    m = build_gen(GEN_FULL, sd, cuda_device)
    x, _, g = make_inputs(200 + T, B, 192, T, 256)
    with torch.no_grad():
        ref = O.generator(sd, x, g)
    got = m(x.to(cuda_device), g=g.to(cuda_device)).cpu()
    assert got.shape == ref.shape == (B, 1, 300 * T)
    assert maxabs(got, ref) <= WAV_TOL


def test_generator_without_g_and_after_remove_weight_norm(cuda_device):
    sd = O.synth_state_dict(gen_shapes(GEN_FULL), 79)
    m = build_gen(GEN_FULL, sd, cuda_device)
    x, _, g = make_inputs(300, 1, 192, 20, 256)
    with torch.no_grad():
        ref_nog = O.generator(sd, x, None)
        ref_g = O.generator(sd, x, g)
    assert maxabs(m(x.to(cuda_device)).cpu(), ref_nog) <= WAV_TOL        # decoder.py:42 skips cond when g is None
    m.remove_weight_norm()                                                # decoder.py:61-65
    assert maxabs(m(x.to(cuda_device), g=g.to(cuda_device)).cpu(), ref_g) <= WAV_TOL


def test_generator_batch_invariance_and_determinism(cuda_device):
    sd = O.synth_state_dict(gen_shapes(GEN_FULL), 80)
    m = build_gen(GEN_FULL, sd, cuda_device)
    x, _, g = make_inputs(301, 3, 192, 40, 256)
    xd, gd = x.to(cuda_device), g.to(cuda_device)
    full = m(xd, g=gd)
    again = m(xd, g=gd)
    assert torch.equal(full, again)                       # run-to-run bit stable (no atomics on the path)
    solo = m(xd[1:2], g=gd[1:2])
    assert torch.equal(full[1:2], solo)                   # utterances are independent


def test_generator_window_matches_oracle_at_bench_size(cuda_device):
    """BASELINE config 3 size (B=16, T=1000): the oracle decodes a 120-frame window of one utterance; away
    from the window edges (receptive field < 40 frames) it must agree with the full-size CUDA run."""
    sd = O.synth_state_dict(gen_shapes(GEN_FULL), 1234)
    m = build_gen(GEN_FULL, sd, cuda_device)
    x, _, g = make_inputs(0, 16, 192, 1000, 256)
    wav = m(x.to(cuda_device), g=g.to(cuda_device))
    assert wav.shape == (16, 1, 300000)
    assert bool(torch.isfinite(wav).all())
    b, t0, t1, rf = 11, 400, 520, 40
    with torch.no_grad():
        ref = O.generator(sd, x[b:b + 1, :, t0:t1], g[b:b + 1])
    got = wav[b:b + 1, :, t0 * 300:t1 * 300].cpu()
    assert maxabs(got[..., rf * 300:-rf * 300], ref[..., rf * 300:-rf * 300]) <= WAV_TOL


def test_flow_at_bench_size_roundtrip_and_window(cuda_device):
    """BASELINE config 2 size (B=16, T=1000): round trip + oracle on two whole utterances."""
    sd = O.synth_state_dict(flow_shapes(FLOW_FULL), 1234)
    m = build_flow(FLOW_FULL, sd, cuda_device)
    x, mask, g = make_inputs(0, 16, 192, 1000, 256)
    xd, md, gd = x.to(cuda_device), mask.to(cuda_device), g.to(cuda_device)
    z = m(xd, md, g=gd, reverse=True)
    back = m(z, md, g=gd, reverse=False)
    assert maxabs(back.cpu(), x) <= Z_TOL
    with torch.no_grad():
        ref = O.flow(sd, x[3:5], mask[3:5], g[3:5], reverse=True)
    assert maxabs(z[3:5].cpu(), ref) <= Z_TOL


# ------------------------------------------------------------------ the whole hot path
def test_prior_sample(cuda_device):
    import ctypes
    from visinger_b200 import _lib
    gen = torch.Generator().manual_seed(9)
    B, C, T = 3, 192, 50
    mu, logs, noise = (torch.randn(B, C, T, generator=gen) for _ in range(3))
    logs = logs * 0.3
    mask = torch.ones(B, 1, T)
    mask[2, :, 31:] = 0
    ref = O.prior_sample(mu, logs, noise, mask)
    d = [t.to(cuda_device).contiguous() for t in (mu, logs, noise, mask)]
    out = torch.empty(B, C, T, device=cuda_device)
    rc = _lib.lib().vsg_prior_sample(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(),
                                     out.data_ptr(), B, C, T, _lib.stream_ptr(cuda_device))
    _lib.check(rc, "vsg_prior_sample")
    assert maxabs(out.cpu(), ref) <= 2e-6
