"""Drop-in for the reference's `modules/visinger/flow.py` (ResidualCouplingBlock and its layers).

Same constructor arguments, same `forward(x, x_mask, g=None, reverse=False)` contract and the same
state-dict keys (`flows.{0,2,..}.pre.{weight,bias}`, `...enc.cond_layer.{bias,weight_g,weight_v}`,
`...enc.in_layers.N.*`, `...enc.res_skip_layers.N.*`, `...post.{weight,bias}`; reference
flow.py:16-31,48-64 and encoder.py:131-165), so `load_state_dict(reference.state_dict())` works.
The arithmetic does not run in PyTorch: forward hands raw device pointers to `vsg_flow_forward`
(include/visinger_b200.h), which runs the fused CUDA kernels.  The nn.Conv1d children below are
parameter containers only (they give the reference's names, shapes and default initialisation).
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.nn.utils import weight_norm, remove_weight_norm

from ... import _lib
from ._packing import PackedModuleMixin


class WaveNet(nn.Module):
    """Parameter layout of reference `WaveNet` (modules/visinger/encoder.py:130-165)."""

    def __init__(self, hidden_channels, kernel_size, dilation_rate, n_layers, gin_channels=0, p_dropout=0):
        super().__init__()
        assert kernel_size % 2 == 1
        self.hidden_channels = hidden_channels
        self.kernel_size = kernel_size
        self.dilation_rate = dilation_rate
        self.n_layers = n_layers
        self.gin_channels = gin_channels
        self.p_dropout = p_dropout
        self.in_layers = nn.ModuleList()
        self.res_skip_layers = nn.ModuleList()
        if gin_channels != 0:
            self.cond_layer = weight_norm(nn.Conv1d(gin_channels, 2 * hidden_channels * n_layers, 1), name="weight")
        for i in range(n_layers):
            dilation = dilation_rate ** i
            padding = int((kernel_size * dilation - dilation) / 2)
            self.in_layers.append(weight_norm(
                nn.Conv1d(hidden_channels, 2 * hidden_channels, kernel_size, dilation=dilation, padding=padding),
                name="weight"))
            rs = 2 * hidden_channels if i < n_layers - 1 else hidden_channels
            self.res_skip_layers.append(weight_norm(nn.Conv1d(hidden_channels, rs, 1), name="weight"))

    def remove_weight_norm(self):
        if self.gin_channels != 0:
            remove_weight_norm(self.cond_layer)
        for l in self.in_layers:
            remove_weight_norm(l)
        for l in self.res_skip_layers:
            remove_weight_norm(l)

    def forward(self, *a, **k):
        raise RuntimeError("visinger_b200 WaveNet holds parameters only; it runs fused inside ResidualCouplingBlock")


class ResidualCouplingLayer(nn.Module):
    """Parameter layout of reference `ResidualCouplingLayer` (flow.py:47-64)."""

    def __init__(self, channels, hidden_channels, kernel_size, dilation_rate, n_layers, p_dropout=0, gin_channels=0,
                 mean_only=False):
        assert channels % 2 == 0, "channels should be divisible by 2"
        super().__init__()
        if not mean_only:
            raise NotImplementedError("only mean_only=True coupling layers exist on the VISinger path (flow.py:29-30)")
        self.channels = channels
        self.hidden_channels = hidden_channels
        self.kernel_size = kernel_size
        self.dilation_rate = dilation_rate
        self.n_layers = n_layers
        self.half_channels = channels // 2
        self.mean_only = mean_only
        self.pre = nn.Conv1d(self.half_channels, hidden_channels, 1)
        self.enc = WaveNet(hidden_channels, kernel_size, dilation_rate, n_layers, p_dropout=p_dropout,
                           gin_channels=gin_channels)
        self.post = nn.Conv1d(hidden_channels, self.half_channels * (2 - mean_only), 1)
        self.post.weight.data.zero_()   # reference flow.py:63-64
        self.post.bias.data.zero_()

    def remove_weight_norm(self):
        self.enc.remove_weight_norm()

    def forward(self, *a, **k):
        raise RuntimeError("visinger_b200 ResidualCouplingLayer runs fused inside ResidualCouplingBlock.forward")


class Flip(nn.Module):
    """Reference flow.py:88-95.  Parameter-free; folded into weight permutations at pack time."""

    def forward(self, x, *args, reverse=False, **kwargs):
        raise RuntimeError("visinger_b200 Flip is folded into the packed weights and never runs on its own")


class ResidualCouplingBlock(PackedModuleMixin, nn.Module):
    """Reference `ResidualCouplingBlock` (modules/visinger/flow.py:15-44) on B200 CUDA kernels.

    `precision`: "fp32" (parity mode: fp32 FFMA kernels; z within 1e-5 of the reference), "bf16x3" (the same tolerance
    on the tcgen05 kernels: three bf16 planes per value) or "bf16" (tcgen05 throughput mode).
    """

    def __init__(self, channels, hidden_channels, kernel_size, dilation_rate, n_layers, n_flows=4, gin_channels=0,
                 precision="fp32"):
        super().__init__()
        self.channels = channels
        self.hidden_channels = hidden_channels
        self.kernel_size = kernel_size
        self.dilation_rate = dilation_rate
        self.n_layers = n_layers
        self.n_flows = n_flows
        self.gin_channels = gin_channels
        self.precision = precision
        self.flows = nn.ModuleList()
        for _ in range(n_flows):
            self.flows.append(ResidualCouplingLayer(channels, hidden_channels, kernel_size, dilation_rate, n_layers,
                                                    gin_channels=gin_channels, mean_only=True))
            self.flows.append(Flip())

    # -- packing ---------------------------------------------------------------------------------
    def _vsg_config(self):
        c = _lib.VsgConfig()
        c.flow_channels, c.flow_hidden, c.flow_kernel_size = self.channels, self.hidden_channels, self.kernel_size
        c.flow_dilation_rate, c.flow_n_layers, c.flow_n_flows = self.dilation_rate, self.n_layers, self.n_flows
        c.flow_gin = self.gin_channels
        return c

    def _vsg_prefixes(self):
        return "", None

    # -- reference API ---------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, x_mask, g=None, reverse=False):
        """x [B, channels, T], x_mask [B, 1, T], g [B, gin, 1] or None -> [B, channels, T] (flow.py:33-40)."""
        _lib.require_cuda(x, "x")
        _lib.require_cuda(x_mask, "x_mask")
        if x.dim() != 3 or x.shape[1] != self.channels:
            raise RuntimeError(f"expected x of shape [B, {self.channels}, T], got {tuple(x.shape)}")
        B, _, T = x.shape
        if x_mask.numel() != B * T:
            raise RuntimeError(f"expected x_mask of shape [B, 1, T] = [{B}, 1, {T}], got {tuple(x_mask.shape)}")
        if self.gin_channels != 0:
            if g is None:
                raise RuntimeError("this flow was built with gin_channels != 0; g is required")
            _lib.require_cuda(g, "g")
            if g.numel() != B * self.gin_channels:
                raise RuntimeError(f"expected g of shape [B, {self.gin_channels}, 1], got {tuple(g.shape)}")
        pack = self._pack()
        prec = _lib.precision_code(self.precision)
        xc, mc = _lib.as_f32c(x), _lib.as_f32c(x_mask)
        gc = _lib.as_f32c(g) if (g is not None and self.gin_channels != 0) else None
        y = torch.empty_like(xc)
        if B == 0 or T == 0:
            return y
        with torch.cuda.device(x.device):
            nbytes = pack.workspace_bytes(B, T, prec)
            ws = _lib.workspace(x.device, nbytes)
            rc = _lib.lib().vsg_flow_forward(pack.handle, xc.data_ptr(), mc.data_ptr(),
                                             gc.data_ptr() if gc is not None else None, y.data_ptr(), B, T,
                                             1 if reverse else 0, prec, ws.data_ptr(), ws.numel(),
                                             _lib.stream_ptr(x.device))
        _lib.check(rc, "vsg_flow_forward")
        return y

    def remove_weight_norm(self):
        # the reference's version (flow.py:42-44) calls a method its layers do not define (SURVEY B-4);
        # this one works.
        for i in range(self.n_flows):
            self.flows[i * 2].remove_weight_norm()
        self.invalidate_pack()
