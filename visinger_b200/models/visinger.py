"""Host-side mirror of the reference's `models/visinger.py` inference branch (lines 105-111).

`HotPath` owns a `ResidualCouplingBlock` (as `.flow`) and a `Generator` (as `.decoder`) -- the
attribute names and state-dict prefixes of the reference `VISinger` module -- packs both into ONE
VsgPack and runs prior sampling -> flow reverse -> decoder through a single `vsg_infer` call
(include/visinger_b200.h), i.e. exactly models/visinger.py:107-111.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib
from ..modules.visinger._packing import PackedModuleMixin
from ..modules.visinger.flow import ResidualCouplingBlock
from ..modules.visinger.decoder import Generator


class HotPath(PackedModuleMixin, nn.Module):
    """models/visinger.py:105-111 as one fused native call."""

    def __init__(self, flow: ResidualCouplingBlock, decoder: Generator, precision: str = "fp32"):
        super().__init__()
        if flow.channels != decoder.initial_channel:
            raise ValueError("flow channels must equal the decoder's initial_channel")
        self.flow = flow
        self.decoder = decoder
        self.precision = precision
        self.last_launches = 0

    @classmethod
    def from_configs(cls, flow_cfg, gen_cfg, flow_sd, gen_sd, device, precision="fp32"):
        flow = ResidualCouplingBlock(flow_cfg["channels"], flow_cfg["hidden"], flow_cfg["kernel_size"],
                                     flow_cfg["dilation_rate"], flow_cfg["n_layers"], n_flows=flow_cfg["n_flows"],
                                     gin_channels=flow_cfg["gin"])
        dec = Generator(gen_cfg["initial_channel"], gen_cfg["resblock"], gen_cfg["rk"], gen_cfg["rd"], gen_cfg["ur"],
                        gen_cfg["uic"], gen_cfg["uk"], gin_channels=gen_cfg["gin"])
        flow.load_state_dict(flow_sd)
        dec.load_state_dict(gen_sd)
        return cls(flow, dec, precision).to(device).eval()

    # -- packing: both modules in one pack, reference checkpoint prefixes ------------------------
    def _vsg_config(self):
        c = self.decoder._vsg_config()
        f = self.flow._vsg_config()
        for name in ("flow_channels", "flow_hidden", "flow_kernel_size", "flow_dilation_rate", "flow_n_layers",
                     "flow_n_flows", "flow_gin"):
            setattr(c, name, getattr(f, name))
        return c

    def _vsg_prefixes(self):
        return "flow.", "decoder."

    # -- models/visinger.py:107-111 --------------------------------------------------------------
    @torch.no_grad()
    def infer(self, mu_p, logs_p, noise, mask, g):
        """(mu_p, logs_p, noise) [B, C, T], mask [B, 1, T], g [B, gin, 1] -> (wav [B, 1, T*hop], z_q [B, C, T])."""
        for n, t in (("mu_p", mu_p), ("logs_p", logs_p), ("noise", noise), ("mask", mask)):
            _lib.require_cuda(t, n)
        B, C, T = mu_p.shape
        if C != self.flow.channels:
            raise RuntimeError(f"expected {self.flow.channels} channels, got {C}")
        if logs_p.shape != mu_p.shape or noise.shape != mu_p.shape or mask.numel() != B * T:
            raise RuntimeError("mu_p, logs_p, noise must share a shape and mask must be [B, 1, T]")
        use_g = self.flow.gin_channels != 0
        if use_g:
            if g is None:
                raise RuntimeError("g is required (gin_channels != 0)")
            _lib.require_cuda(g, "g")
        pack = self._pack()
        prec = _lib.precision_code(self.precision)
        a = [_lib.as_f32c(t) for t in (mu_p, logs_p, noise, mask)]
        gc = _lib.as_f32c(g) if use_g else None
        dev = mu_p.device
        wav = torch.empty(B, 1, T * self.decoder.hop_size, dtype=torch.float32, device=dev)
        z_q = torch.empty(B, C, T, dtype=torch.float32, device=dev)
        if B == 0 or T == 0:
            return wav, z_q
        with torch.cuda.device(dev):
            ws = _lib.workspace(dev, pack.workspace_bytes(B, T, prec))
            rc = _lib.lib().vsg_infer(pack.handle, a[0].data_ptr(), a[1].data_ptr(), a[2].data_ptr(), a[3].data_ptr(),
                                      gc.data_ptr() if use_g else None, wav.data_ptr(), z_q.data_ptr(), B, T, prec,
                                      ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
        _lib.check(rc, "vsg_infer")
        self.last_launches = _lib.last_launch_count()
        return wav, z_q

    # -- CUDA-graph replay of the whole path -----------------------------------------------------
    @torch.no_grad()
    def graph(self, B: int, T: int, device=None) -> "HotPathGraph":
        """Capture prior sampling -> flow reverse -> decoder for a fixed (B, T) into one CUDA graph.

        The run calls of the C ABI neither allocate nor synchronise, so the ~300-1200 kernel launches of one
        pass (more with L2-resident batch tiling) replay from a single cudaGraphLaunch.  Inputs are written into
        the graph's static buffers (`.mu_p, .logs_p, .noise, .mask, .g`), outputs read from `.wav, .z_q`."""
        dev = torch.device(device) if device is not None else next(self.parameters()).device
        key = (B, T, _lib.precision_code(self.precision), str(dev))
        cache = self.__dict__.setdefault("_graphs", {})
        if key not in cache:
            cache[key] = HotPathGraph(self, B, T, dev)
        return cache[key]

    @torch.no_grad()
    def decode(self, z, g):
        """Generator only, through the shared pack (used by bench.py for the decoder roofline)."""
        _lib.require_cuda(z, "z")
        B, _, T = z.shape
        pack = self._pack()
        prec = _lib.precision_code(self.precision)
        zc = _lib.as_f32c(z)
        use_g = self.decoder.gin_channels != 0
        gc = _lib.as_f32c(g) if use_g else None
        wav = torch.empty(B, 1, T * self.decoder.hop_size, dtype=torch.float32, device=z.device)
        with torch.cuda.device(z.device):
            ws = _lib.workspace(z.device, pack.workspace_bytes(B, T, prec))
            rc = _lib.lib().vsg_generator_forward(pack.handle, zc.data_ptr(), gc.data_ptr() if use_g else None,
                                                  wav.data_ptr(), B, T, prec, ws.data_ptr(), ws.numel(),
                                                  _lib.stream_ptr(z.device))
        _lib.check(rc, "vsg_generator_forward")
        self.last_launches = _lib.last_launch_count()
        return wav


class HotPathGraph:
    """A captured (B, T)-shaped pass of `HotPath`: static input/output/workspace buffers + one torch.cuda.CUDAGraph."""

    def __init__(self, hp: HotPath, B: int, T: int, device: torch.device):
        C, gin = hp.flow.channels, hp.flow.gin_channels
        f32 = dict(dtype=torch.float32, device=device)
        self.mu_p = torch.zeros(B, C, T, **f32)
        self.logs_p = torch.zeros(B, C, T, **f32)
        self.noise = torch.zeros(B, C, T, **f32)
        self.mask = torch.ones(B, 1, T, **f32)
        self.g = torch.zeros(B, max(gin, 1), 1, **f32)
        self.wav = torch.empty(B, 1, T * hp.decoder.hop_size, **f32)
        self.z_q = torch.empty(B, C, T, **f32)
        pack = hp._pack()
        prec = _lib.precision_code(hp.precision)
        self._ws = torch.empty(pack.workspace_bytes(B, T, prec), dtype=torch.uint8, device=device)  # private: pointers are baked
        self._pack = pack
        L = _lib.lib()

        def run():
            rc = L.vsg_infer(pack.handle, self.mu_p.data_ptr(), self.logs_p.data_ptr(), self.noise.data_ptr(),
                             self.mask.data_ptr(), self.g.data_ptr() if gin else None, self.wav.data_ptr(),
                             self.z_q.data_ptr(), B, T, prec, self._ws.data_ptr(), self._ws.numel(),
                             _lib.stream_ptr(device))
            _lib.check(rc, "vsg_infer")

        with torch.cuda.device(device):
            side = torch.cuda.Stream(device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):      # warm-up outside capture (one-time function attributes, lazy module load)
                run()
            torch.cuda.current_stream(device).wait_stream(side)
            torch.cuda.synchronize(device)
            self.launches = _lib.last_launch_count()
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                run()

    def replay(self):
        self._graph.replay()
        return self.wav, self.z_q

    def __call__(self, mu_p, logs_p, noise, mask, g=None):
        self.mu_p.copy_(mu_p, non_blocking=True)
        self.logs_p.copy_(logs_p, non_blocking=True)
        self.noise.copy_(noise, non_blocking=True)
        self.mask.copy_(mask, non_blocking=True)
        if g is not None:
            self.g.copy_(g, non_blocking=True)
        return self.replay()
