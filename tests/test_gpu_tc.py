"""GPU: the bf16 tcgen05 path.  (1) per-layer parity of the tensor-core Conv1d against F.conv1d on the same
bf16-rounded operands (fp32 accumulate: only the summation order differs), (2) Generator in bf16 mode against
the fp32 oracle, reported as relative L2 (BASELINE.json: bf16 mode is reported, not gated at 1e-4)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import visinger_oracle as O
from helpers import gen_shapes, make_inputs, build_gen, maxabs, GEN_FULL

pytestmark = pytest.mark.gpu

CONV_TOL = 2e-3


CASES = [
    (64, 64, 3, 1, 1, 128), (256, 256, 3, 1, 2, 300), (256, 256, 11, 5, 2, 300), (128, 128, 7, 3, 2, 700),
    (64, 64, 11, 1, 3, 129), (32, 32, 3, 3, 2, 300), (16, 16, 7, 5, 2, 300), (192, 512, 7, 1, 2, 100),
    (512, 256, 3, 1, 2, 200), (96, 192, 1, 1, 2, 300), (192, 384, 5, 1, 2, 300), (16, 16, 3, 1, 1, 20000),
    (32, 32, 11, 5, 1, 5000), (128, 128, 3, 5, 1, 1000),
]


def _case(cin, cout, k, dil, B, L):
    gen = torch.Generator().manual_seed(cin + cout + k + dil)
    x = torch.randn(B, L, cin, generator=gen).to(torch.bfloat16)
    w = (torch.randn(cout, cin, k, generator=gen) / (cin * k) ** 0.5).to(torch.bfloat16).float()
    b = torch.randn(cout, generator=gen) * 0.1
    ref = F.conv1d(x.float().transpose(1, 2).double(), w.double(), b.double(), dilation=dil,
                   padding=(k - 1) * dil // 2).transpose(1, 2)
    return x, w, b, ref, gen


@pytest.mark.parametrize("flags", [0, 1, 3 | (1 << 4), 3 | (2 << 4), 3],
                         ids=["reload", "halo", "halo+resident,mb1", "halo+resident,mb<=2", "halo+resident,mb<=4"])
@pytest.mark.parametrize("cin,cout,k,dil,B,L", CASES)
def test_tc_conv1d_matches_fconv1d(cuda_device, cin, cout, k, dil, B, L, flags):
    """A-operand feeding modes: per-tap reload, halo (row-shifted UMMA descriptors), halo + resident weights, and
    1 / 2 / 4 blocks of 128 rows per tile (flags bits 4..: cap on blocks per tile)."""
    from visinger_b200 import _lib
    x, w, b, ref, _ = _case(cin, cout, k, dil, B, L)
    got = _lib.debug_conv1d_bf16(x.to(cuda_device).contiguous(), w, b, dil, flags=flags).cpu()
    assert maxabs(got, ref) <= CONV_TOL


@pytest.mark.parametrize("mbcap", [1, 2, 4])
@pytest.mark.parametrize("cin,cout,k,dil,B,L", [(256, 256, 3, 1, 2, 300), (128, 128, 7, 3, 1, 333), (64, 64, 11, 5, 2, 257),
                                                (32, 32, 7, 1, 2, 1000), (16, 16, 3, 1, 3, 4100), (16, 16, 11, 5, 1, 70),
                                                (64, 64, 3, 1, 2, 513), (32, 32, 11, 5, 1, 511)])
def test_tc_conv1d_fused_epilogue(cuda_device, cin, cout, k, dil, B, L, mbcap):
    """Residual + running-sum adds (TMA-loaded), scale, and the two bf16 outputs (TMA-stored): what the
    reference does as `x = xt + x`, `xs += ...`, `x = xs / 3`, `F.leaky_relu` (decoder.py:48-54,93-102)."""
    from visinger_b200 import _lib
    x, w, b, ref, gen = _case(cin, cout, k, dil, B, L)
    add0 = torch.randn(B, L, cout, generator=gen).to(torch.bfloat16)
    add1 = torch.randn(B, L, cout, generator=gen).to(torch.bfloat16)
    want = (ref + add0.double() + add1.double()) / 3.0
    d = cuda_device
    out, raw, act = _lib.debug_conv1d_bf16(x.to(d).contiguous(), w, b, dil, flags=3 | (mbcap << 4), add0=add0.to(d).contiguous(),
                                           add1=add1.to(d).contiguous(), scale=1.0 / 3.0, want_bf16=True)
    assert maxabs(out.cpu(), want) <= CONV_TOL
    assert maxabs(raw.cpu().double(), want) <= 2e-2          # bf16 rounding of O(1) values
    assert maxabs(act.cpu().double(), F.leaky_relu(want, 0.1)) <= 2e-2
    # bf16 outputs are exactly the rounded fp32 result (the TMA store path moves bytes, it does not compute)
    assert torch.equal(raw.cpu(), out.cpu().to(torch.bfloat16))


@pytest.mark.parametrize("cin,cout,k,dil,B,L", [(256, 256, 3, 1, 2, 300), (32, 32, 7, 3, 2, 1000), (16, 16, 11, 5, 1, 700)])
def test_tc_conv1d_residual_from_activated_stream(cuda_device, cin, cout, k, dil, B, L):
    """Plain-bf16 decoder keeps ONE copy of the resblock stream, a = leaky_relu(x); `x = xt + x` (decoder.py:102)
    recovers x from it in the epilogue.  bf16(0.1 x) * 10 carries the same relative rounding as bf16(x)."""
    from visinger_b200 import _lib
    x, w, b, ref, gen = _case(cin, cout, k, dil, B, L)
    res = torch.randn(B, L, cout, generator=gen)
    a = F.leaky_relu(res, 0.1).to(torch.bfloat16)
    # what the kernel reconstructs: min(a, a * 10) on packed bf16 pairs (one more bf16 rounding for negative values)
    resid = torch.minimum(a, a * torch.tensor(10.0, dtype=torch.bfloat16)).float()
    assert maxabs(resid, res) <= 4e-2
    d = cuda_device
    out = _lib.debug_conv1d_bf16(x.to(d).contiguous(), w, b, dil, flags=3 | 8, add0=a.to(d).contiguous())
    assert maxabs(out.cpu(), ref + resid.double()) <= CONV_TOL


@pytest.mark.parametrize("B,T", [(1, 7), (2, 64), (3, 150)])
def test_generator_bf16_vs_oracle(cuda_device, B, T):
    sd = O.synth_state_dict(gen_shapes(GEN_FULL), 1234)
    m = build_gen(GEN_FULL, sd, cuda_device, precision="bf16")
    x, _, g = make_inputs(500 + T, B, 192, T, 256)
    with torch.no_grad():
        ref = O.generator(sd, x, g)
    got = m(x.to(cuda_device), g=g.to(cuda_device)).cpu()
    assert got.shape == ref.shape
    rel = float((got - ref).norm() / ref.norm())
    print(f"bf16 generator B={B} T={T}: rel-L2 {rel:.3e}, max-abs {maxabs(got, ref):.3e}, |ref|max {float(ref.abs().max()):.3e}")
    assert rel <= 5.7e-3 and maxabs(got, ref) <= 5.5e-4      # 1.5 x the measured floor (3.8e-3 / 3.5e-4; |ref|max 0.038)
    m32 = build_gen(GEN_FULL, sd, cuda_device, precision="fp32")
    assert maxabs(m32(x.to(cuda_device), g=g.to(cuda_device)).cpu(), ref) <= 1e-4


def test_generator_bf16_deterministic_and_batch_invariant(cuda_device):
    sd = O.synth_state_dict(gen_shapes(GEN_FULL), 99)
    m = build_gen(GEN_FULL, sd, cuda_device, precision="bf16")
    x, _, g = make_inputs(900, 3, 192, 50, 256)
    xd, gd = x.to(cuda_device), g.to(cuda_device)
    a, b = m(xd, g=gd), m(xd, g=gd)
    assert torch.equal(a, b)
    assert torch.equal(a[2:3], m(xd[2:3], g=gd[2:3]))


def test_generator_bf16_full_length_fused_equals_unfused(cuda_device):
    """BASELINE-size utterances (T = 1000 frames = 300 000 samples: thousands of 256- / 512-row tiles per utterance in the
    fused-pair stages) are beyond what the CPU oracle finishes in seconds, so the full-length check is a property: the
    fused ResBlock-pair kernels, the merged upsamplers, the N-split and the signature-specialised images must reproduce
    what the plain one-kernel-per-conv path computes.  The two paths accumulate in different orders, every bf16 rounding
    that flips is re-amplified by the following layers, so they differ by about sqrt(2) x the bf16-vs-fp32 error of
    either (3.7e-3) -- a tiling / halo / indexing bug shows up as O(1).  Together with the small-size oracle tests above
    this ties the full-length run to the reference."""
    from visinger_b200 import _lib
    sd = O.synth_state_dict(gen_shapes(GEN_FULL), 99)
    m = build_gen(GEN_FULL, sd, cuda_device, precision="bf16")
    x, _, g = make_inputs(901, 2, 192, 1000, 256)
    xd, gd = x.to(cuda_device), g.to(cuda_device)
    try:
        fast = m(xd, g=gd).clone()
        # bit 9 no fused pairs, bit 10 one launch per polyphase, bit 11 no N split, bit 13 generic epilogue images
        _lib.set_tc_options(halo_mode=1 | 512 | 1024 | 2048 | 8192)
        plain = m(xd, g=gd).clone()
    finally:
        _lib.set_tc_options(halo_mode=1)
    assert fast.shape == (2, 1, 300000) and bool(torch.isfinite(fast).all())
    rel = float((fast - plain).norm() / plain.norm())
    print(f"full-length bf16 generator, fused vs plain path: rel-L2 {rel:.3e}, max-abs {maxabs(fast, plain):.3e}")
    assert rel <= 1e-2          # measured 5.6e-3 = sqrt(2) x 3.7e-3 (two independent bf16 roundings of the same function)
    # the fp32 FFMA path (pinned to the oracle at small sizes, different kernels and layout) at the same full length
    ref32 = build_gen(GEN_FULL, sd, cuda_device, precision="fp32")(xd, g=gd)
    rel16 = float((fast - ref32).norm() / ref32.norm())
    rel16p = float((plain - ref32).norm() / ref32.norm())
    x3 = build_gen(GEN_FULL, sd, cuda_device, precision="bf16x3")(xd, g=gd)
    print(f"full length vs fp32 path: bf16 rel-L2 {rel16:.3e} (plain path {rel16p:.3e}); bf16x3 max-abs {maxabs(x3, ref32):.3e}")
    assert rel16 <= 1e-2 and rel16 <= 1.5 * rel16p   # bf16 noise floor of this network; the fast path adds none of its own
    assert maxabs(x3, ref32) <= 1e-4          # north_star's fp32 tolerance, at full length


@pytest.mark.parametrize("B,T", [(1, 1), (1, 3), (3, 17), (5, 86), (2, 171), (1, 1707), (33, 64)])
def test_generator_bf16_odd_shapes_vs_fp32_path(cuda_device, B, T):
    """Ragged tile counts everywhere (partial 128 / 256 / 512-row tiles, utterances shorter than one fused-pair tile so the
    un-fused fallback runs, more utterances than SMs' worth of tiles): the bf16 path stays at its noise floor against the
    fp32 path (itself pinned to the oracle) and is deterministic."""
    sd = O.synth_state_dict(gen_shapes(GEN_FULL), 5)
    m16 = build_gen(GEN_FULL, sd, cuda_device, precision="bf16")
    m32 = build_gen(GEN_FULL, sd, cuda_device, precision="fp32")
    x, _, g = make_inputs(B * 1000 + T, B, 192, T, 256)
    xd, gd = x.to(cuda_device), g.to(cuda_device)
    fast, ref = m16(xd, g=gd), m32(xd, g=gd)
    assert bool(torch.isfinite(fast).all()) and torch.equal(fast, m16(xd, g=gd))
    rel = float((fast - ref).norm() / ref.norm())
    assert rel <= 1.2e-2, rel        # 8.1e-3 for these weights at every shape (tools/stress_shapes.py)


@pytest.mark.parametrize("B,T,lengths", [(1, 1, None), (2, 300, [300, 211]), (3, 130, [130, 128, 5])])
def test_flow_bf16_vs_oracle(cuda_device, B, T, lengths):
    """bf16 tensor-core flow (gate / residual-skip / coupling epilogues) against the fp32 oracle.  bf16 storage of
    the coupling state bounds the accuracy; BASELINE.json asks for it to be reported, the 1e-5 gate is fp32-mode."""
    from helpers import flow_shapes, build_flow, FLOW_FULL
    sd = O.synth_state_dict(flow_shapes(FLOW_FULL), 1234)
    m = build_flow(FLOW_FULL, sd, cuda_device, precision="bf16")
    x, mask, g = make_inputs(40 + T, B, 192, T, 256, lengths)
    x = x * mask
    with torch.no_grad():
        ref_rev = O.flow(sd, x, mask, g, reverse=True)
        ref_fwd = O.flow(sd, x, mask, g, reverse=False)
    d = cuda_device
    rev = m(x.to(d), mask.to(d), g=g.to(d), reverse=True).cpu()
    fwd = m(x.to(d), mask.to(d), g=g.to(d), reverse=False).cpu()
    rel = float((rev - ref_rev).norm() / ref_rev.norm())
    print(f"bf16 flow B={B} T={T}: reverse rel-L2 {rel:.3e} max-abs {maxabs(rev, ref_rev):.3e}; "
          f"forward max-abs {maxabs(fwd, ref_fwd):.3e}; |z|max {float(ref_rev.abs().max()):.2f}")
    # 1.5 x the measured floors (rel-L2 3.2e-3; max-abs 3.0e-2 reverse, 4.1e-2 forward at |z|max ~4.7)
    assert rel <= 4.8e-3 and maxabs(rev, ref_rev) <= 4.5e-2 and maxabs(fwd, ref_fwd) <= 6.2e-2
    if lengths is not None:
        for b, n in enumerate(lengths):
            assert float(rev[b, :, n:].abs().max()) == 0.0 if n < T else True


def test_hot_path_bf16_and_fp32(cuda_device):
    """vsg_infer (models/visinger.py:107-111 in one call) in both modes against the oracle."""
    from helpers import flow_shapes, FLOW_FULL
    from visinger_b200.models.visinger import HotPath
    fsd = O.synth_state_dict(flow_shapes(FLOW_FULL), 1234)
    gsd = O.synth_state_dict(gen_shapes(GEN_FULL), 1234)
    sd = {"flow." + k: v for k, v in fsd.items()}
    sd.update({"decoder." + k: v for k, v in gsd.items()})
    B, T = 3, 90
    mu, mask, g = make_inputs(77, B, 192, T, 256, [90, 64, 33])
    gen = torch.Generator().manual_seed(5)
    logs = 0.3 * torch.randn(B, 192, T, generator=gen) - 1.0
    noise = torch.randn(B, 192, T, generator=gen)
    with torch.no_grad():
        wav_ref, z_ref = O.infer_hot_path(sd, mu, logs, noise, mask, g)
    d = cuda_device
    args = [t.to(d) for t in (mu, logs, noise, mask, g)]
    hp32 = HotPath.from_configs(FLOW_FULL, GEN_FULL, fsd, gsd, d, precision="fp32")
    wav, z = hp32.infer(*args)
    assert maxabs(z.cpu(), z_ref) <= 1e-5 and maxabs(wav.cpu().squeeze(1), wav_ref) <= 1e-4
    hp16 = HotPath.from_configs(FLOW_FULL, GEN_FULL, fsd, gsd, d, precision="bf16")
    wav16, z16 = hp16.infer(*args)
    rel = float((wav16.cpu().squeeze(1) - wav_ref).norm() / wav_ref.norm())
    print(f"bf16 hot path: wav rel-L2 {rel:.3e}, z max-abs {maxabs(z16.cpu(), z_ref):.3e}")
    assert rel <= 5.7e-3 and maxabs(z16.cpu(), z_ref) <= 4.5e-2      # 1.5 x floor (3.79e-3 / 2.9e-2)


def test_hot_path_pipeline_matches_direct_calls(cuda_device):
    """HotPath.pipeline (upload / run / download on three streams, double-buffered graph slots) must return, for a
    stream of DIFFERENT host-resident requests, exactly what hp.infer returns for each of them (fp32 mode: the
    same kernels, so bit-identical) -- i.e. no slot is overwritten while a neighbouring request still uses it."""
    from helpers import flow_shapes, FLOW_FULL
    from visinger_b200.models.visinger import HotPath
    fsd = O.synth_state_dict(flow_shapes(FLOW_FULL), 1234)
    gsd = O.synth_state_dict(gen_shapes(GEN_FULL), 1234)
    d = cuda_device
    hp = HotPath.from_configs(FLOW_FULL, GEN_FULL, fsd, gsd, d, precision="fp32")
    B, T, n_req = 2, 40, 5
    pipe = hp.pipeline(B, T, d)
    reqs = []
    for r in range(n_req):
        mu, mask, g = make_inputs(300 + r, B, 192, T, 256, [40, 17 + r])
        gen = torch.Generator().manual_seed(r)
        logs = 0.3 * torch.randn(B, 192, T, generator=gen) - 1.0
        noise = torch.randn(B, 192, T, generator=gen)
        reqs.append([t.contiguous().pin_memory() for t in (mu, logs, noise, mask, g)])
    tickets, got = [], {}
    for r, host in enumerate(reqs):
        tickets.append(pipe.submit(*host))
        if r >= 1:                                   # collect with one request of lag, like a server would
            got[r - 1] = pipe.result(tickets[r - 1]).clone()
    got[n_req - 1] = pipe.result(tickets[-1]).clone()
    with pytest.raises(RuntimeError):
        pipe.result(tickets[0])                      # its slot has been reused
    for r, host in enumerate(reqs):
        wav, _ = hp.infer(*[t.to(d) for t in host])
        assert torch.equal(wav.cpu().view(B, -1), got[r]), f"request {r}"


@pytest.mark.parametrize("cin,cout,k,dil,B,L", [(256, 256, 3, 1, 2, 300), (128, 128, 11, 5, 1, 700), (64, 64, 7, 3, 2, 513),
                                                (32, 32, 11, 1, 1, 1000), (16, 16, 3, 5, 2, 4100), (192, 512, 7, 1, 1, 100)])
def test_tc_conv1d_split_bf16(cuda_device, cin, cout, k, dil, B, L):
    """Split-bf16 (bf16x3) mode: fp32 operands carried as (hi, lo) bf16 pairs, three MMAs per product.  Against an
    fp64 convolution of the UN-rounded fp32 operands the error must be ~1e-5, not the ~4e-3 of plain bf16."""
    from visinger_b200 import _lib
    gen = torch.Generator().manual_seed(cin * 3 + cout + k + dil)
    x = torch.randn(B, L, cin, generator=gen)
    w = torch.randn(cout, cin, k, generator=gen) / (cin * k) ** 0.5
    b = torch.randn(cout, generator=gen) * 0.1
    add0 = torch.randn(B, L, cout, generator=gen)
    ref = F.conv1d(x.transpose(1, 2).double(), w.double(), b.double(), dilation=dil, padding=(k - 1) * dil // 2).transpose(1, 2)
    want = (ref + add0.double()) * 0.5
    d = cuda_device
    out, raw, act = _lib.debug_conv1d_bf16(_lib.split_bf16(x).to(d), w, b, dil, flags=3 | 4,
                                           add0=_lib.split_bf16(add0).to(d), scale=0.5, want_bf16=True)
    assert maxabs(out.cpu(), want) <= 5e-5
    assert maxabs(_lib.merge_bf16(raw.cpu()), want) <= 1e-4
    assert maxabs(_lib.merge_bf16(act.cpu()), F.leaky_relu(want, 0.1)) <= 1e-4


@pytest.mark.parametrize("B,T", [(1, 7), (2, 64)])
def test_generator_bf16x3_meets_fp32_tolerance(cuda_device, B, T):
    """The split-bf16 tensor-core decoder must stay inside the fp32-mode waveform tolerance (max-abs <= 1e-4)."""
    sd = O.synth_state_dict(gen_shapes(GEN_FULL), 1234)
    m = build_gen(GEN_FULL, sd, cuda_device, precision="bf16x3")
    x, _, g = make_inputs(500 + T, B, 192, T, 256)
    with torch.no_grad():
        ref = O.generator(sd, x, g)
    got = m(x.to(cuda_device), g=g.to(cuda_device)).cpu()
    err = maxabs(got, ref)
    print(f"bf16x3 generator B={B} T={T}: max-abs {err:.3e}, rel-L2 {float((got - ref).norm() / ref.norm()):.3e}")
    assert err <= 1e-4


@pytest.mark.parametrize("C,k,d1,B,L", [(16, 3, 1, 2, 1000), (16, 3, 5, 1, 777), (16, 7, 3, 2, 512), (16, 11, 5, 1, 2049),
                                        (32, 3, 3, 2, 640), (32, 7, 1, 1, 1500), (32, 11, 5, 2, 256), (32, 11, 1, 1, 4100)])
def test_fused_resblock_pair_kernel(cuda_device, C, k, d1, B, L):
    """One (c1, c2) step of ResBlock1 (decoder.py:93-102) in a single kernel: conv1 -> +b1 -> leaky_relu -> (bf16, in
    shared memory) -> conv2 -> +b2 + residual + running sum -> x scale.  The reference below rounds the intermediate
    to bf16 exactly where the kernel does."""
    from visinger_b200 import _lib
    gen = torch.Generator().manual_seed(C * 7 + k * 3 + d1)
    x = torch.randn(B, L, C, generator=gen).to(torch.bfloat16)
    xa = F.leaky_relu(x.float(), 0.1).to(torch.bfloat16)
    w1 = (torch.randn(C, C, k, generator=gen) / (C * k) ** 0.5).to(torch.bfloat16).float()
    w2 = (torch.randn(C, C, k, generator=gen) / (C * k) ** 0.5).to(torch.bfloat16).float()
    b1, b2 = torch.randn(C, generator=gen) * 0.1, torch.randn(C, generator=gen) * 0.1
    add1 = torch.randn(B, L, C, generator=gen).to(torch.bfloat16)
    t = F.conv1d(xa.float().transpose(1, 2).double(), w1.double(), b1.double(), dilation=d1, padding=(k - 1) * d1 // 2)
    t = F.leaky_relu(t, 0.1).float().to(torch.bfloat16)
    y = F.conv1d(t.double(), w2.double(), b2.double(), padding=(k - 1) // 2).transpose(1, 2)
    want = (y + x.double() + add1.double()) / 3.0
    d = cuda_device
    two_adds = not (C == 32 and k == 11)      # 2 adds + 2 outputs at C=32, k=11 exceeds shared memory (never occurs:
    if not two_adds:                           # the decoder's last pair has 2 adds but 1 output)
        want = (y + x.double()) / 3.0
    out, raw, act = _lib.debug_pair_bf16(xa.to(d).contiguous(), w1, b1, w2, b2, d1, add0=x.to(d).contiguous(),
                                         add1=add1.to(d).contiguous() if two_adds else None, scale=1.0 / 3.0)
    err = (out.cpu().double() - want).abs()
    # an intermediate element that sits on a bf16 rounding boundary may round the other way (fp32 vs fp64 accumulation):
    # bound the worst case loosely and the typical case tightly
    assert float(err.max()) <= 1e-2 and float(err.mean()) <= 2e-4
    assert torch.equal(raw.cpu(), out.cpu().to(torch.bfloat16))
    assert maxabs(act.cpu().double(), F.leaky_relu(out.cpu().double(), 0.1)) <= 2e-2
