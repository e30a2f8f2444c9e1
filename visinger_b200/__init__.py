"""visinger_b200 -- B200-native (sm_100a) implementation of the VISinger inference hot path.

Mirrors the reference's module surface for that path:

    visinger_b200.modules.visinger.flow.ResidualCouplingBlock   (reference modules/visinger/flow.py:15)
    visinger_b200.modules.visinger.decoder.Generator            (reference modules/visinger/decoder.py:13)
    visinger_b200.models.visinger.VISinger                      (reference models/visinger.py:18)

All three run on hand-written CUDA kernels behind the C ABI of include/visinger_b200.h.
"""
from ._lib import build, lib, last_launch_count  # noqa: F401
from .modules.visinger.flow import ResidualCouplingBlock  # noqa: F401
from .modules.visinger.decoder import Generator  # noqa: F401

__all__ = ["build", "lib", "ResidualCouplingBlock", "Generator", "last_launch_count"]
