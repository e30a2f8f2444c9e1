"""length regulator + frame positions (`vsg_length_regulate`), the fused frame-prior head
(`vsg_frame_prior_forward`: encoder -> proj -> split -> prior sampling) and `vsg_infer_zp`.

The oracle side is the reference's own arithmetic: `expand_states` (models/commons/align_ops.py:22-26) and
`SinusoidalPositionalEmbedding` (modules/rel_transformer.py:45-100) restated below, `O.rel_encoder` (pinned to the
reference fixtures in test_relenc.py) + `proj` + `O.prior_sample`."""
import pytest
import torch
import torch.nn.functional as F

from oracle import visinger_oracle as O
from helpers import maxabs, FLOW_FULL, GEN_FULL, flow_shapes, gen_shapes, make_inputs


def _expand_states(h, mel2token):
    """models/commons/align_ops.py:22-26"""
    h = F.pad(h, [0, 0, 1, 0])
    return torch.gather(h, 1, mel2token[..., None].repeat([1, 1, h.shape[-1]]))


def _sin_table(n, dim):
    """modules/rel_transformer.py:60-77 get_embedding(padding_idx=0)"""
    import math
    half = dim // 2
    emb = math.log(10000) / (half - 1)
    emb = torch.exp(torch.arange(half, dtype=torch.float) * -emb)
    emb = torch.arange(n, dtype=torch.float).unsqueeze(1) * emb.unsqueeze(0)
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=1).view(n, -1)
    emb[0, :] = 0
    return emb


def _regulate_ref(enc, mel2ph, table):
    """models/visinger.py:76-82: expand, then add the table rows of make_positions(channel 0) (rel_transformer.py:78-88)."""
    x = _expand_states(enc.transpose(1, 2), mel2ph)                      # [B, T, H]
    if table is not None:
        keep = x[..., 0].ne(0).int()
        pos = (torch.cumsum(keep, dim=1).type_as(keep) * keep).long()
        x = x + table.index_select(0, pos.view(-1)).view(x.shape[0], x.shape[1], -1)
    return x.transpose(1, 2)


def _mel2ph(B, T_ph, T, gen, lengths):
    m = torch.zeros(B, T, dtype=torch.long)
    for b in range(B):
        n_tok = int(torch.randint(3, T_ph + 1, (1,), generator=gen))
        durs = torch.randint(1, 2 * lengths[b] // n_tok + 2, (n_tok,), generator=gen)
        idx = torch.repeat_interleave(torch.arange(1, n_tok + 1), durs)[:lengths[b]]
        m[b, :idx.numel()] = idx
    return m


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,T_ph,T,use_pos", [(3, 192, 40, 1000, True), (2, 32, 7, 65, True), (2, 48, 12, 513, False),
                                                (1, 192, 90, 1280, True)])
def test_length_regulate_matches_the_reference_formula(cuda_device, B, H, T_ph, T, use_pos):
    from visinger_b200 import _lib
    gen = torch.Generator().manual_seed(B * 1000 + T)
    enc = torch.randn(B, H, T_ph, generator=gen)
    enc[0, 0, 2] = 0.0                                                   # a token whose channel 0 is exactly zero: its frames
    lengths = [T] + [max(1, T - 37 * (b + 1)) for b in range(B - 1)]    # get no position (the reference's data-dependent rule)
    mel2ph = _mel2ph(B, T_ph, T, gen, lengths)
    table = _sin_table(T + 1, H) if use_pos else None
    want = _regulate_ref(enc, mel2ph, table)
    got = _lib.length_regulate(enc.to(cuda_device), mel2ph.to(cuda_device), table.to(cuda_device) if use_pos else None)
    assert got.shape == want.shape
    assert torch.equal(got.cpu(), want)                                  # a gather and one fp32 add: bit-exact
    assert float(got.cpu()[1, :, lengths[1]:].abs().max()) == 0.0 if B > 1 else True


def _frame_prior_case(seed, B, T, lengths, hidden=192, filt=768, n_layers=4):
    shapes = {"encoder." + k: v for k, v in O.rel_encoder_param_shapes(hidden, filt, 2, n_layers, 9, 4, 1).items()}
    shapes["proj.weight"] = (2 * hidden, hidden, 1)
    shapes["proj.bias"] = (2 * hidden,)
    sd = O.synth_rel_encoder_state_dict(shapes, seed)
    gen = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, hidden, T, generator=gen)
    g = torch.randn(B, 1, T, generator=gen)
    noise = torch.randn(B, hidden, T, generator=gen)
    mask = torch.ones(B, 1, T)
    for b, n in enumerate(lengths):
        mask[b, :, n:] = 0
    with torch.no_grad():
        h = O.rel_encoder(sd, x * mask, mask, g * mask, n_heads=2, n_layers=n_layers, kernel_size=9, window=4, prefix="encoder.")
        stats = F.conv1d(h, sd["proj.weight"], sd["proj.bias"]) * mask                 # modules/visinger/encoder.py:71
        mu, logs = torch.split(stats, hidden, dim=1)
        z = O.prior_sample(mu, logs, noise, mask)                                        # models/visinger.py:107
    return sd, x * mask, mask, g * mask, noise, z, mu, logs


def _build_frame_prior(sd, device, precision, hidden=192, filt=768, n_layers=4):
    from visinger_b200.modules.visinger.encoder import FramePriorNetwork
    m = FramePriorNetwork(hidden, filt, 2, n_layers, 9, gin_channels=1, p_dropout=0.1)
    m.load_state_dict(sd, strict=True)
    m = m.to(device).eval()
    m.precision = precision
    return m


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_frame_prior_head_vs_oracle(cuda_device, precision):
    sd, x, mask, g, noise, z, mu, logs = _frame_prior_case(41, 2, 300, [300, 211])
    m = _build_frame_prior(sd, cuda_device, precision)
    zq, mu_g, logs_g = m.sample(x.to(cuda_device), mask.to(cuda_device), g.to(cuda_device), noise.to(cuda_device))
    if precision == "fp32":
        errs = [maxabs(zq.cpu(), z), maxabs(mu_g.cpu(), mu), maxabs(logs_g.cpu(), logs)]
        print(f"frame prior head fp32: z / mu / logs max-abs {errs[0]:.3e} {errs[1]:.3e} {errs[2]:.3e}")
        assert errs[0] <= 5e-5 and max(errs[1:]) <= 2e-5            # z carries noise * exp(logs): |z| up to ~6
        # the module's reference-shaped forward (mu_p, logs_p) agrees with the fused head
        with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):   # (proj is a cuDNN conv there)
            mu2, logs2 = m(x.to(cuda_device), mask.to(cuda_device), g.to(cuda_device))
        assert maxabs(mu2, mu_g) <= 2e-5 and maxabs(logs2, logs_g) <= 2e-5
    else:
        rel = float((mu_g.cpu() - mu).norm() / mu.norm())
        relz = float((zq.cpu() - z).norm() / z.norm())
        print(f"frame prior head bf16: mu rel-L2 {rel:.3e}, z rel-L2 {relz:.3e}, logs max-abs {maxabs(logs_g.cpu(), logs):.3e}")
        assert rel <= 8.5e-3 and relz <= 6e-3                       # 1.5x the measured 5.6e-3 / 4.0e-3
    assert float(zq.cpu()[1, :, 211:].abs().max()) == 0.0
    z2, _, _ = m.sample(x.to(cuda_device), mask.to(cuda_device), g.to(cuda_device), noise.to(cuda_device))
    assert torch.equal(zq, z2)


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("bf16x3", 2e-5), ("bf16", 0.0)])
def test_infer_zp_equals_infer(cuda_device, precision, tol):
    """vsg_infer_zp(z_p) must reproduce vsg_infer(mu_p, logs_p, noise) when z_p is that call's own prior sample."""
    from visinger_b200.models.visinger import HotPath
    fsd = O.synth_state_dict(flow_shapes(FLOW_FULL), 77)
    gsd = O.synth_state_dict(gen_shapes(GEN_FULL), 78)
    hp = HotPath.from_configs(FLOW_FULL, GEN_FULL, fsd, gsd, cuda_device, precision=precision)
    x, mask, g = make_inputs(9, 2, 192, 40, 256, [40, 29])
    gen = torch.Generator().manual_seed(3)
    logs = 0.3 * torch.randn(2, 192, 40, generator=gen) - 1.0
    noise = torch.randn(2, 192, 40, generator=gen)
    a = [t.to(cuda_device) for t in (x, logs, noise, mask, g)]
    wav, zq = hp.infer(*a)
    from visinger_b200 import _lib
    z_p = torch.empty_like(a[0])
    _lib.check(_lib.lib().vsg_prior_sample(a[0].data_ptr(), a[1].data_ptr(), a[2].data_ptr(), a[3].data_ptr(), z_p.data_ptr(),
                                           2, 192, 40, _lib.stream_ptr(cuda_device)), "vsg_prior_sample")
    wav2, zq2 = hp.infer_zp(z_p, a[3], a[4])
    assert torch.equal(wav, wav2) and torch.equal(zq, zq2)
