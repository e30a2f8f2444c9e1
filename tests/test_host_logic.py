"""CPU: host-side mirror of the reference interface (constructor signatures, state-dict contract,
loud failure without CUDA)."""
import pytest
import torch

from oracle import visinger_oracle as O
from helpers import FLOW_FULL, GEN_FULL, flow_shapes, gen_shapes


def test_flow_state_dict_contract():
    from visinger_b200 import ResidualCouplingBlock
    m = ResidualCouplingBlock(192, 192, 5, 1, 4, gin_channels=256)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == flow_shapes(FLOW_FULL)
    # reference zero-initialises `post` (flow.py:63-64)
    assert float(m.flows[0].post.weight.abs().max()) == 0.0
    sd = O.synth_state_dict(flow_shapes(FLOW_FULL), 3)
    m.load_state_dict(sd)   # strict
    assert len(m.flows) == 8 and not list(m.flows[1].parameters())


def test_generator_state_dict_contract_and_remove_weight_norm():
    from visinger_b200 import Generator
    m = Generator(192, "1", [3, 7, 11], [[1, 3, 5]] * 3, [5, 5, 3, 2, 2], 512, [11, 11, 7, 4, 4], gin_channels=256)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == gen_shapes(GEN_FULL)
    assert m.hop_size == 300
    # ConvTranspose1d weight-norm is over dim 0 = C_in
    assert got["ups.0.weight_g"] == (512, 1, 1) and got["ups.0.weight_v"] == (512, 256, 11)
    m.load_state_dict(O.synth_state_dict(gen_shapes(GEN_FULL), 4))
    m.remove_weight_norm()
    keys = set(m.state_dict().keys())
    assert "ups.0.weight" in keys and "ups.0.weight_g" not in keys
    assert "resblocks.0.convs1.0.weight" in keys


def test_resblock2_layout():
    from visinger_b200 import Generator
    cfg = dict(initial_channel=8, resblock="2", rk=[3, 5], rd=[[1, 3], [1, 3]], ur=[4, 2], uic=16, uk=[8, 4], gin=0)
    m = Generator(8, "2", cfg["rk"], cfg["rd"], cfg["ur"], 16, cfg["uk"], gin_channels=0)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == gen_shapes(cfg)


def test_cpu_tensors_fail_loudly_no_fallback():
    from visinger_b200 import ResidualCouplingBlock, Generator
    f = ResidualCouplingBlock(8, 16, 5, 1, 2, gin_channels=0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        f(torch.zeros(1, 8, 10), torch.ones(1, 1, 10))
    g = Generator(8, "1", [3], [[1, 3, 5]], [2], 16, [4], gin_channels=0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        g(torch.zeros(1, 8, 10))


def test_product_package_does_not_import_the_oracle():
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dp, _, files in os.walk(os.path.join(root, "visinger_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("no CPU or", ""), f"{f} mentions the oracle"


def test_random_init_matches_reference_parameter_layout():
    """HotPath.random_init (what bench.py's own arm uses instead of oracle weights) builds the reference architecture:
    same parameter names and shapes as the oracle's shape tables (= the reference's state dict), non-zero `post` layers."""
    from oracle import visinger_oracle as O
    from visinger_b200.configs import VISINGER_FLOW as F, VISINGER_GENERATOR as G
    from visinger_b200.models.visinger import HotPath
    hp = HotPath.random_init(F, G, "cpu", precision="bf16", seed=7)
    sd = hp.state_dict()
    want = {"flow." + k: v for k, v in O.flow_param_shapes(F["channels"], F["hidden"], F["kernel_size"], F["n_layers"],
                                                            F["n_flows"], F["gin"]).items()}
    want.update({"decoder." + k: v for k, v in O.generator_param_shapes(G["initial_channel"], G["resblock"], G["rk"], G["rd"],
                                                                        G["ur"], G["uic"], G["uk"], G["gin"]).items()})
    assert set(sd) == set(want)
    for k, shp in want.items():
        assert tuple(sd[k].shape) == tuple(shp), k
    assert all(float(sd[k].abs().max()) > 0 for k in sd if ".post.weight" in k)
    again = HotPath.random_init(F, G, "cpu", precision="bf16", seed=7).state_dict()
    assert all(torch.equal(sd[k], again[k]) for k in sd)


def test_three_plane_plans_exist_for_every_flow_convolution():
    """Host-only planner (no GPU): the chunk table of a three-plane convolution has four sweeps over the channel chunks
    (csrc/run_tc.cu::launch_conv_tc), which must fit the kernel's table and shared memory for every convolution of the
    flow at config/models/visinger.yaml sizes (pre 96 -> 192 / 384, in_layer 192 -> 384 k5, res_skip 192 -> 384 / 192,
    post 192 -> 96) in both split modes, with and without residual operands."""
    import ctypes
    from visinger_b200 import _lib
    L = _lib.lib()
    L.vsg_debug_plan.argtypes = [ctypes.c_int32] * 9
    L.vsg_debug_plan.restype = ctypes.c_int
    for planes_code in (0, 2):
        for cin, cout, k, n_adds in ((96, 192, 1, 0), (96, 384, 1, 0), (192, 384, 5, 0), (192, 384, 1, 1), (192, 192, 1, 1),
                                    (192, 96, 1, 1)):
            for B, T in ((16, 1000), (1, 1), (3, 130)):
                rc = L.vsg_debug_plan(cin, cout, k, 1, B, T, n_adds, 1, planes_code)
                assert rc == 0, (planes_code, cin, cout, k, _lib.last_error_message() if hasattr(_lib, "last_error_message") else rc)


def test_rowpacked_resblock_plans():
    """Host-only planner of the row-packed whole-ResBlock1 kernel (csrc/run_tc.cu::rp_plan): at the model's C = 32 / C = 16
    stages (decoder.py:91-104 with kernel sizes 3 / 7 / 11, dilations 1 / 3 / 5) every variant must fit its shared-memory
    and tensor-memory budget, cover an utterance with whole tiles, and recompute exactly the halo the six convolutions reach;
    shapes the kernel does not take must be refused, not mis-planned."""
    import ctypes
    from visinger_b200 import _lib
    L = _lib.lib()
    L.vsg_debug_rp_plan.argtypes = [ctypes.c_int32] * 3 + [ctypes.POINTER(ctypes.c_int32), ctypes.c_int32, ctypes.c_int32,
                                                            ctypes.POINTER(ctypes.c_int32)]
    L.vsg_debug_rp_plan.restype = ctypes.c_int
    dil = (ctypes.c_int32 * 3)(1, 3, 5)
    out = (ctypes.c_int32 * 8)()
    smem_max = 227 * 1024
    for C, rate in ((32, 150), (16, 300)):
        S = 64 // C
        for k in (3, 7, 11):
            for T in (1000, 37, 2):
                Lx = rate * T
                for variant in (0, 1, 2):
                    rc = L.vsg_debug_rp_plan(C, k, 3, dil, Lx, variant, out)
                    assert rc == 0, (C, k, T, variant)
                    mb, H, V, ring, smem, packed, tiles, tmem = list(out)
                    halo = (k - 1) // 2 * (1 + 1 + 3 + 1 + 5 + 1)
                    assert H == (halo + S - 1) // S * S and V == 128 * S * mb - 2 * H and V > 0
                    assert tiles * V >= Lx and (tiles - 1) * V < Lx
                    assert 1 <= mb <= (4 if variant == 0 else 2) and ring >= 2
                    assert smem <= (smem_max if variant < 2 else (smem_max - 2048) // 2)
                    assert tmem >= 2 * 64 * mb and tmem <= (512 if variant < 2 else 256)
                    assert packed == 0b101011          # conv1 of the dilation-1 pair and every conv2 in the Toeplitz form
    # refused: an odd number of time steps at C = 32 (rows hold two), even kernel sizes, C = 64 in the split-bf16 form
    assert L.vsg_debug_rp_plan(32, 3, 3, dil, 301, 0, out) != 0
    assert L.vsg_debug_rp_plan(16, 4, 3, dil, 3000, 0, out) != 0
    assert L.vsg_debug_rp_plan(64, 3, 3, dil, 3000, 1, out) != 0
