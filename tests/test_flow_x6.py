"""The flow at north_star's fp32 tolerance (z max-abs <= 1e-5) ON THE TENSOR CORES: precision "bf16x3" runs
ResidualCouplingBlock (modules/visinger/flow.py:33-40) through the tcgen05 kernels with three bf16 planes per value
(hi, mid, lo = the fp32 value exactly), six plane products per MMA product, small products first (the tensor pipe's
fp32 accumulation truncates: tools/acc_probe.py, tests/emulate_tc_accumulate.py).  Checked against the CPU oracle, the
reference-made golden and, at bench size, the fp32 CUDA-core path."""
import pytest
import torch

import visinger_b200
from oracle import visinger_oracle as O
from helpers import load_npz, flow_shapes, make_inputs, build_flow, maxabs, FLOW_FULL

pytestmark = pytest.mark.gpu

Z_TOL = 1e-5


def _run(m, x, mask, g, d, reverse):
    return m(x.to(d), mask.to(d), g=g.to(d) if g is not None else None, reverse=reverse).cpu()


@pytest.mark.parametrize("B,T,lengths", [(1, 1, None), (2, 300, [300, 211]), (3, 130, [130, 128, 5]), (1, 1000, None)])
def test_flow_bf16x3_tensor_core_vs_oracle(cuda_device, B, T, lengths):
    sd = O.synth_state_dict(flow_shapes(FLOW_FULL), 1234)
    x, mask, g = make_inputs(40 + T, B, 192, T, 256, lengths)
    x = x * mask
    with torch.no_grad():
        ref_rev = O.flow(sd, x, mask, g, reverse=True)
        ref_fwd = O.flow(sd, x, mask, g, reverse=False)
    m16 = build_flow(FLOW_FULL, sd, cuda_device, precision="bf16")
    _run(m16, x, mask, g, cuda_device, True)
    n_tc = visinger_b200.last_launch_count()
    m = build_flow(FLOW_FULL, sd, cuda_device, precision="bf16x3")
    rev = _run(m, x, mask, g, cuda_device, True)
    assert visinger_b200.last_launch_count() == n_tc, "bf16x3 flow did not run the tensor-core launch sequence"
    fwd = _run(m, x, mask, g, cuda_device, False)
    e_rev, e_fwd = maxabs(rev, ref_rev), maxabs(fwd, ref_fwd)
    print(f"bf16x3 flow (tcgen05, 3 planes) B={B} T={T}: reverse max-abs {e_rev:.3e}, forward {e_fwd:.3e} "
          f"(|z|max {float(ref_rev.abs().max()):.2f})")
    assert e_rev <= Z_TOL and e_fwd <= Z_TOL
    assert torch.equal(rev, _run(m, x, mask, g, cuda_device, True))          # run-to-run bit-stable
    if lengths is not None:
        for b, n in enumerate(lengths):
            if n < T:
                assert float(rev[b, :, n:].abs().max()) == 0.0


def test_flow_bf16x3_full_config_golden(cuda_device):
    """The reference-made fixture (tests/golden/make_golden.py imported the unmodified reference modules)."""
    z = load_npz("full_flow")
    sd = O.synth_state_dict(flow_shapes(FLOW_FULL), int(z["seed"]))
    m = build_flow(FLOW_FULL, sd, cuda_device, precision="bf16x3")
    x, mask, g = make_inputs(int(z["seed"]) + 1, int(z["B"]), 192, int(z["T"]), 256, z["lengths"].tolist())
    x = x * mask
    st = int(z["slice_t"])
    rev = _run(m, x, mask, g, cuda_device, True)
    fwd = _run(m, x, mask, g, cuda_device, False)
    assert maxabs(rev[:, :, ::st], torch.from_numpy(z["z_rev"])) <= Z_TOL
    assert maxabs(fwd[:, :, ::st], torch.from_numpy(z["z_fwd"])) <= Z_TOL
    assert maxabs(rev[:, :, ::st], torch.from_numpy(z["z_rev64"])) <= Z_TOL


def test_flow_bf16x3_dilated_odd_flows(cuda_device):
    """dilation_rate > 1, an odd number of flows (one net channel flip) and no speaker condition on the three-plane path."""
    cfg = dict(channels=64, hidden=32, kernel_size=5, dilation_rate=2, n_layers=3, n_flows=3, gin=0)
    sd = O.synth_state_dict(flow_shapes(cfg), 99)
    x, mask, g = make_inputs(5, 2, 64, 257, 0, [257, 100])
    x = x * mask
    kw = dict(channels=64, hidden=32, kernel_size=5, dilation_rate=2, n_layers=3, n_flows=3)
    with torch.no_grad():
        ref_rev = O.flow(sd, x, mask, None, reverse=True, **kw)
        ref_fwd = O.flow(sd, x, mask, None, reverse=False, **kw)
    m = build_flow(cfg, sd, cuda_device, precision="bf16x3")
    rev, fwd = _run(m, x, mask, None, cuda_device, True), _run(m, x, mask, None, cuda_device, False)
    m32 = build_flow(cfg, sd, cuda_device, precision="fp32")
    _run(m32, x, mask, None, cuda_device, True)
    n32 = visinger_b200.last_launch_count()
    _run(m, x, mask, None, cuda_device, True)
    assert visinger_b200.last_launch_count() != n32, "expected the tensor-core launch sequence"
    assert maxabs(rev, ref_rev) <= Z_TOL and maxabs(fwd, ref_fwd) <= Z_TOL


def test_flow_bf16x3_bench_size_vs_fp32_path(cuda_device):
    """BASELINE.json configs[1] (B = 16 x T = 1000): the tensor-core flow against the fp32 CUDA-core path, plus the
    round trip forward(reverse(x)) = x."""
    sd = O.synth_state_dict(flow_shapes(FLOW_FULL), 1234)
    x, mask, g = make_inputs(0, 16, 192, 1000, 256, [1000] * 12 + [777, 512, 130, 1])
    x = x * mask
    d = cuda_device
    xd, md, gd = x.to(d), mask.to(d), g.to(d)
    m = build_flow(FLOW_FULL, sd, d, precision="bf16x3")
    m32 = build_flow(FLOW_FULL, sd, d, precision="fp32")
    z = m(xd, md, g=gd, reverse=True)
    z32 = m32(xd, md, g=gd, reverse=True)
    back = m(z * md, md, g=gd, reverse=False)
    e, rt = maxabs(z, z32), maxabs(back * md, xd)
    print(f"bf16x3 flow at B16 x T1000: vs the fp32 path max-abs {e:.3e}, round trip {rt:.3e}")
    assert e <= Z_TOL and rt <= 2 * Z_TOL
