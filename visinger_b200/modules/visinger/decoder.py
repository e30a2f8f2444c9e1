"""Drop-in for the reference's `modules/visinger/decoder.py` (HiFi-GAN `Generator`).

Same constructor, `forward(x, g=None) -> [B, 1, T*hop]`, `remove_weight_norm()` and state-dict keys
(`conv_pre.*`, `ups.N.{bias,weight_g,weight_v}`, `resblocks.N.convs{1,2}.M.*`, `conv_post.weight`,
`cond.*`; reference decoder.py:14-38,69-89,114-122).  forward() calls `vsg_generator_forward`
(include/visinger_b200.h); the nn.Conv1d / nn.ConvTranspose1d children are parameter containers.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.nn.utils import weight_norm, remove_weight_norm

from ... import _lib
from ._packing import PackedModuleMixin

LRELU_SLOPE = 0.1


def get_padding(kernel_size, dilation=1):
    return int((kernel_size * dilation - dilation) / 2)


class ResBlock1(nn.Module):
    """Parameter layout of reference ResBlock1 (decoder.py:68-89)."""

    def __init__(self, channels, kernel_size=3, dilation=(1, 3, 5)):
        super().__init__()
        self.kernel_size, self.dilation = kernel_size, tuple(dilation)
        self.convs1 = nn.ModuleList([
            weight_norm(nn.Conv1d(channels, channels, kernel_size, 1, dilation=d, padding=get_padding(kernel_size, d)))
            for d in dilation])
        self.convs2 = nn.ModuleList([
            weight_norm(nn.Conv1d(channels, channels, kernel_size, 1, dilation=1, padding=get_padding(kernel_size, 1)))
            for _ in dilation])

    def remove_weight_norm(self):
        for l in list(self.convs1) + list(self.convs2):
            remove_weight_norm(l)

    def forward(self, *a, **k):
        raise RuntimeError("visinger_b200 ResBlock1 holds parameters only; it runs fused inside Generator.forward")


class ResBlock2(nn.Module):
    """Parameter layout of reference ResBlock2 (decoder.py:113-122)."""

    def __init__(self, channels, kernel_size=3, dilation=(1, 3)):
        super().__init__()
        self.kernel_size, self.dilation = kernel_size, tuple(dilation)
        self.convs = nn.ModuleList([
            weight_norm(nn.Conv1d(channels, channels, kernel_size, 1, dilation=d, padding=get_padding(kernel_size, d)))
            for d in dilation])

    def remove_weight_norm(self):
        for l in self.convs:
            remove_weight_norm(l)

    def forward(self, *a, **k):
        raise RuntimeError("visinger_b200 ResBlock2 holds parameters only; it runs fused inside Generator.forward")


class Generator(PackedModuleMixin, nn.Module):
    """Reference `Generator` (modules/visinger/decoder.py:13-65) on B200 CUDA kernels.

    `precision`: "fp32" (parity mode, waveform within 1e-4 of the reference) or "bf16" (tcgen05 mode).
    """

    def __init__(self, initial_channel, resblock, resblock_kernel_sizes, resblock_dilation_sizes, upsample_rates,
                 upsample_initial_channel, upsample_kernel_sizes, gin_channels=0, precision="fp32"):
        super().__init__()
        self.num_kernels = len(resblock_kernel_sizes)
        self.num_upsamples = len(upsample_rates)
        self.initial_channel = initial_channel
        self.resblock_type = 1 if resblock == "1" else 2
        self.resblock_kernel_sizes = [int(k) for k in resblock_kernel_sizes]
        self.resblock_dilation_sizes = [[int(d) for d in ds] for ds in resblock_dilation_sizes]
        self.upsample_rates = [int(u) for u in upsample_rates]
        self.upsample_kernel_sizes = [int(k) for k in upsample_kernel_sizes]
        self.upsample_initial_channel = upsample_initial_channel
        self.gin_channels = gin_channels
        self.precision = precision
        self.hop_size = 1
        for u in self.upsample_rates:
            self.hop_size *= u

        self.conv_pre = nn.Conv1d(initial_channel, upsample_initial_channel, 7, 1, padding=3)
        block = ResBlock1 if self.resblock_type == 1 else ResBlock2
        self.ups = nn.ModuleList()
        for i, (u, k) in enumerate(zip(self.upsample_rates, self.upsample_kernel_sizes)):
            self.ups.append(weight_norm(nn.ConvTranspose1d(upsample_initial_channel // (2 ** i),
                                                           upsample_initial_channel // (2 ** (i + 1)), k, u,
                                                           padding=(k - u) // 2)))
        self.resblocks = nn.ModuleList()
        ch = upsample_initial_channel
        for i in range(len(self.ups)):
            ch = upsample_initial_channel // (2 ** (i + 1))
            for k, d in zip(self.resblock_kernel_sizes, self.resblock_dilation_sizes):
                self.resblocks.append(block(ch, k, d))
        self.conv_post = nn.Conv1d(ch, 1, 7, 1, padding=3, bias=False)
        if gin_channels != 0:
            self.cond = nn.Conv1d(gin_channels, upsample_initial_channel, 1)

    # -- packing ---------------------------------------------------------------------------------
    def _vsg_config(self):
        c = _lib.VsgConfig()
        c.dec_initial_channel = self.initial_channel
        c.dec_resblock = self.resblock_type
        c.dec_n_kernels = self.num_kernels
        if self.num_kernels > _lib.VSG_MAX_RESBLOCK_KERNELS or self.num_upsamples > _lib.VSG_MAX_UPS:
            raise RuntimeError("too many resblock kernels / upsampling stages for the packed configuration")
        for j, (k, ds) in enumerate(zip(self.resblock_kernel_sizes, self.resblock_dilation_sizes)):
            if len(ds) > _lib.VSG_MAX_RESBLOCK_DILATIONS:
                raise RuntimeError("too many dilations per resblock")
            c.dec_resblock_kernel_sizes[j] = k
            c.dec_n_dilations[j] = len(ds)
            for q, d in enumerate(ds):
                c.dec_resblock_dilations[j][q] = d
        c.dec_n_ups = self.num_upsamples
        for i, (u, k) in enumerate(zip(self.upsample_rates, self.upsample_kernel_sizes)):
            c.dec_upsample_rates[i] = u
            c.dec_upsample_kernel_sizes[i] = k
        c.dec_upsample_initial_channel = self.upsample_initial_channel
        c.dec_gin = self.gin_channels
        return c

    def _vsg_prefixes(self):
        return None, ""

    # -- reference API ---------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, g=None):
        """x [B, initial_channel, T], g [B, gin, 1] or None -> waveform [B, 1, T * hop] (decoder.py:40-59)."""
        _lib.require_cuda(x, "x")
        if x.dim() != 3 or x.shape[1] != self.initial_channel:
            raise RuntimeError(f"expected x of shape [B, {self.initial_channel}, T], got {tuple(x.shape)}")
        B, _, T = x.shape
        use_g = g is not None and self.gin_channels != 0
        if use_g:
            _lib.require_cuda(g, "g")
            if g.numel() != B * self.gin_channels:
                raise RuntimeError(f"expected g of shape [B, {self.gin_channels}, 1], got {tuple(g.shape)}")
        pack = self._pack_for(use_g)
        prec = _lib.precision_code(self.precision)
        xc = _lib.as_f32c(x)
        gc = _lib.as_f32c(g) if use_g else None
        wav = torch.empty(B, 1, T * self.hop_size, dtype=torch.float32, device=x.device)
        if B == 0 or T == 0:
            return wav
        with torch.cuda.device(x.device):
            nbytes = pack.workspace_bytes(B, T, prec)
            ws = _lib.workspace(x.device, nbytes)
            rc = _lib.lib().vsg_generator_forward(pack.handle, xc.data_ptr(), gc.data_ptr() if use_g else None,
                                                  wav.data_ptr(), B, T, prec, ws.data_ptr(), ws.numel(),
                                                  _lib.stream_ptr(x.device))
        _lib.check(rc, "vsg_generator_forward")
        return wav

    def _pack_for(self, use_g):
        # The reference skips `cond` when g is None (decoder.py:42) even if the layer exists; a pack
        # built with dec_gin = 0 reproduces that without a second code path in the kernels.
        if use_g or self.gin_channels == 0:
            return self._pack()
        if getattr(self, "_vsg_pack_nog", None) is None or self._vsg_key_nog != self._current_key():
            cfg = self._vsg_config()
            cfg.dec_gin = 0
            dev = next(self.parameters()).device
            self._vsg_pack_nog = _lib.Pack(cfg, dict(self.state_dict()), None, "", dev)
            self._vsg_key_nog = self._current_key()
        return self._vsg_pack_nog

    def _current_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def remove_weight_norm(self):
        for l in self.ups:
            remove_weight_norm(l)
        for l in self.resblocks:
            l.remove_weight_norm()
        self.invalidate_pack()
