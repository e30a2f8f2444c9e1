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


@pytest.mark.parametrize("cin,cout,k,dil,B,L", [
    (64, 64, 3, 1, 1, 128), (256, 256, 3, 1, 2, 300), (256, 256, 11, 5, 2, 300), (128, 128, 7, 3, 2, 700),
    (64, 64, 11, 1, 3, 129), (32, 32, 3, 3, 2, 300), (16, 16, 7, 5, 2, 300), (192, 512, 7, 1, 2, 100),
    (512, 256, 3, 1, 2, 200), (96, 192, 1, 1, 2, 300), (192, 384, 5, 1, 2, 300), (16, 16, 3, 1, 1, 20000),
])
def test_tc_conv1d_matches_fconv1d(cuda_device, cin, cout, k, dil, B, L):
    from visinger_b200 import _lib
    gen = torch.Generator().manual_seed(cin + cout + k + dil)
    x = torch.randn(B, L, cin, generator=gen).to(torch.bfloat16)
    w = (torch.randn(cout, cin, k, generator=gen) / (cin * k) ** 0.5).to(torch.bfloat16).float()
    b = torch.randn(cout, generator=gen) * 0.1
    ref = F.conv1d(x.float().transpose(1, 2).double(), w.double(), b.double(), dilation=dil,
                   padding=(k - 1) * dil // 2).transpose(1, 2)
    got = _lib.debug_conv1d_bf16(x.to(cuda_device).contiguous(), w, b, dil, flags=0).cpu()
    assert maxabs(got, ref) <= CONV_TOL


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
    assert rel <= 3e-2
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
