"""PyTorch mirrors of the prior-side modules of the reference's `modules/visinger/encoder.py`:
TextEncoder (:14-55), FramePriorNetwork (:58-73) and a parameter-only PosteriorEncoder (:76-101).

Not on the B200 hot path ("next" rows); they exist so that the `VISinger` mirror is a complete
`forward(infer=True)` drop-in that loads reference checkpoints.  `WaveNet` lives in `.flow` (it is part of the
hot path there)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import _lib
from ..rel_transformer import RelativeEncoder, SinusoidalPositionalEmbedding
from .flow import WaveNet

DEFAULT_MAX_TARGET_POSITIONS = 2000


def Embedding(num_embeddings, embedding_dim, padding_idx=None):
    """reference modules/commons/utils.py:71-76"""
    m = nn.Embedding(num_embeddings, embedding_dim, padding_idx=padding_idx)
    nn.init.normal_(m.weight, mean=0.0, std=embedding_dim ** -0.5)
    if padding_idx is not None:
        nn.init.constant_(m.weight[padding_idx], 0)
    return m


def expand_states(h, mel2token):
    """Length regulator (reference models/commons/align_ops.py:22-26): frame t copies token mel2token[t] (1-based);
    0 selects an all-zero row.  h [B, T_ph, H], mel2token [B, T] -> [B, T, H]."""
    h = F.pad(h, [0, 0, 1, 0])
    return torch.gather(h, 1, mel2token[..., None].expand(-1, -1, h.shape[-1]))


class TextEncoder(nn.Module):
    def __init__(self, ph_dict_size, note_pitch_size, note_dur_size, hidden_channels, filter_channels, n_heads, n_layers,
                 kernel_size, p_dropout, use_pos_embed=False):
        super().__init__()
        self.use_pos_embed = use_pos_embed
        self.ph_emb = Embedding(ph_dict_size, hidden_channels)
        self.pitch_emb = Embedding(note_pitch_size, hidden_channels)
        self.dur_emb = Embedding(note_dur_size, hidden_channels)
        self.embed_scale = math.sqrt(hidden_channels)
        self.linear = nn.Linear(hidden_channels * 3, hidden_channels)
        self.text_encoder = RelativeEncoder(hidden_channels, filter_channels, n_heads, n_layers, kernel_size=kernel_size,
                                            p_dropout=p_dropout)
        if use_pos_embed:
            self.embed_positions = SinusoidalPositionalEmbedding(hidden_channels, 0, init_size=DEFAULT_MAX_TARGET_POSITIONS)

    def forward(self, text_tokens, pitch_tokens, dur_tokens, mel2ph):
        enc = self.encode_tokens(text_tokens, pitch_tokens, dur_tokens)
        return expand_states(enc.transpose(1, 2), mel2ph).transpose(1, 2)    # [B, H, T]

    def encode_tokens(self, text_tokens, pitch_tokens, dur_tokens):
        """Everything before the length regulator: [B, T_ph] tokens -> phoneme-rate states [B, H, T_ph]."""
        nonpad = (text_tokens > 0).float().unsqueeze(1)                      # [B, 1, T_ph]
        emb = torch.cat([self.ph_emb(text_tokens), self.pitch_emb(pitch_tokens), self.dur_emb(dur_tokens)], 2)
        tok = self.linear(emb * self.embed_scale) * nonpad.transpose(1, 2)   # [B, T_ph, H]
        if self.use_pos_embed:
            # Reference quirk kept bit-for-bit (encoder.py:50-53): seq_len is passed as tok.shape[2] (= H, not T_ph), so
            # the [B*T_ph, H] table lookup is *viewed* as [B, H, T_ph] and then transposed -- a scrambled embedding.
            pos = self.embed_positions(tok.shape[0], tok.shape[2], tok[..., 0])
            tok = tok + pos.transpose(1, 2)
        tok = tok * nonpad.transpose(1, 2)
        return self.text_encoder(tok.transpose(1, 2), nonpad)                # [B, H, T_ph]


class FramePriorNetwork(nn.Module):
    def __init__(self, hidden_channels, filter_channels, n_heads, n_layers, kernel_size, gin_channels, p_dropout):
        super().__init__()
        self.hidden_channels = hidden_channels
        self.encoder = RelativeEncoder(hidden_channels, filter_channels, n_heads, n_layers=n_layers, kernel_size=kernel_size,
                                       gin_channels=gin_channels, p_dropout=p_dropout)
        self.proj = nn.Conv1d(hidden_channels, hidden_channels * 2, 1)

    def forward(self, x, x_mask, g=None):
        # The reference transposes g [B, 1, T] -> [B, T, 1] here (encoder.py:68-69) and then crashes in Conv1d(1 -> H)
        # (SURVEY.md Appendix B-1).  The evident intent -- condition every frame on its log-f0 -- is implemented.
        out = self.proj(self.encoder(x, x_mask, g)) * x_mask
        return torch.split(out, self.hidden_channels, dim=1)

    precision = "fp32"

    @torch.no_grad()
    def sample(self, x, x_mask, g=None, noise=None):
        """Native fused head (vsg_frame_prior_forward): encoder -> proj -> split -> prior sampling
        (models/visinger.py:107) in one library call.  Returns (z_p, mu_p, logs_p), each [B, H, T]."""
        _lib.require_cuda(x, "x")
        B, H, T = x.shape
        if noise is None:
            noise = torch.randn(B, H, T, device=x.device, dtype=torch.float32)
        enc = self.encoder
        params = list(self.parameters())
        key = (str(x.device),) + tuple((p.data_ptr(), p._version) for p in params)
        if getattr(self, "_vsg_pack", None) is None or self._vsg_key != key:
            c = _lib.VsgRelEncConfig(enc.hidden_channels, enc.filter_channels, enc.n_heads, enc.n_layers, enc.kernel_size,
                                     enc.window_size, enc.gin_channels or 0)
            self._vsg_pack = _lib.RelEncPack(c, dict(self.state_dict()), "", x.device, frame_prior=True)
            self._vsg_key = key
        pack = self._vsg_pack
        prec = _lib.precision_code(self.precision)
        xc, mc, nc = _lib.as_f32c(x), _lib.as_f32c(x_mask), _lib.as_f32c(noise)
        gc = _lib.as_f32c(g) if g is not None else None
        if gc is not None and tuple(gc.shape) != (B, 1, T):
            raise RuntimeError(f"expected g of shape [B, 1, T] = {(B, 1, T)}, got {tuple(gc.shape)}")
        stats = torch.empty(B, 2 * H, T, device=x.device, dtype=torch.float32)
        z = torch.empty(B, H, T, device=x.device, dtype=torch.float32)
        if B and T:
            with torch.cuda.device(x.device):
                nbytes = int(_lib.lib().vsg_frame_prior_workspace_bytes(pack.handle, B, T, prec))
                ws = _lib.workspace(x.device, nbytes)
                rc = _lib.lib().vsg_frame_prior_forward(pack.handle, xc.data_ptr(), mc.data_ptr(),
                                                        gc.data_ptr() if gc is not None else None, nc.data_ptr(),
                                                        stats.data_ptr(), z.data_ptr(), B, T, prec, ws.data_ptr(), ws.numel(),
                                                        _lib.stream_ptr(x.device))
            _lib.check(rc, "vsg_frame_prior_forward")
        mu_p, logs_p = torch.split(stats, H, dim=1)
        return z, mu_p, logs_p


class PosteriorEncoder(nn.Module):
    """Reference `PosteriorEncoder` (modules/visinger/encoder.py:76-101) on the B200 kernels: `pre` (1x1) -> WaveNet ->
    `proj` (1x1) -> reparameterised sample, one `vsg_posterior_forward` call.  Same
    constructor, state-dict keys and `forward(x, nonpadding, g)` -> `(z_q, mu_q, logs_q)` contract; `noise` (defaults to
    `torch.randn_like(mu_q)`, encoder.py:97) can be injected for parity tests.  `precision`: "fp32" (FFMA parity
    kernels) or "bf16" (tcgen05)."""

    def __init__(self, in_channels, out_channels, hidden_channels, kernel_size, dilation_rate, n_layers, gin_channels,
                 precision="fp32"):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.hidden_channels = hidden_channels
        self.kernel_size = kernel_size
        self.dilation_rate = dilation_rate
        self.n_layers = n_layers
        self.gin_channels = gin_channels
        self.precision = precision
        self.pre = nn.Conv1d(in_channels, hidden_channels, 1)
        self.enc = WaveNet(hidden_channels, kernel_size, dilation_rate, n_layers, gin_channels=gin_channels)
        self.proj = nn.Conv1d(hidden_channels, out_channels * 2, 1)
        self._vsg_pack = None
        self._vsg_key = None

    def _pack(self):
        params = list(self.parameters())
        dev = params[0].device
        key = (str(dev),) + tuple((p.data_ptr(), p._version) for p in params)
        if self._vsg_pack is None or self._vsg_key != key:
            c = _lib.VsgEncConfig(self.in_channels, self.out_channels, self.hidden_channels, self.kernel_size,
                                  self.dilation_rate, self.n_layers, self.gin_channels)
            self._vsg_pack = _lib.EncPack(c, dict(self.state_dict()), "", dev)
            self._vsg_key = key
        return self._vsg_pack

    @torch.no_grad()
    def forward(self, x, nonpadding, g=None, noise=None):
        """x [B, in_channels, T], nonpadding [B, 1, T], g [B, gin, 1] -> (z_q, mu_q, logs_q), each [B, out_channels, T]."""
        _lib.require_cuda(x, "x")
        _lib.require_cuda(nonpadding, "nonpadding")
        if x.dim() != 3 or x.shape[1] != self.in_channels:
            raise RuntimeError(f"expected x of shape [B, {self.in_channels}, T], got {tuple(x.shape)}")
        B, _, T = x.shape
        if nonpadding.numel() != B * T:
            raise RuntimeError(f"expected nonpadding of shape [B, 1, T] = [{B}, 1, {T}], got {tuple(nonpadding.shape)}")
        if self.gin_channels != 0:
            if g is None:
                raise RuntimeError("this encoder was built with gin_channels != 0; g is required")
            _lib.require_cuda(g, "g")
            if g.numel() != B * self.gin_channels:
                raise RuntimeError(f"expected g of shape [B, {self.gin_channels}, 1], got {tuple(g.shape)}")
        if noise is None:
            noise = torch.randn(B, self.out_channels, T, device=x.device, dtype=torch.float32)
        elif tuple(noise.shape) != (B, self.out_channels, T):
            raise RuntimeError(f"expected noise of shape {(B, self.out_channels, T)}, got {tuple(noise.shape)}")
        pack = self._pack()
        prec = _lib.precision_code(self.precision)
        xc, mc, nc = _lib.as_f32c(x), _lib.as_f32c(nonpadding), _lib.as_f32c(noise)
        gc = _lib.as_f32c(g) if (g is not None and self.gin_channels != 0) else None
        z = torch.empty(B, self.out_channels, T, device=x.device, dtype=torch.float32)
        stats = torch.empty(B, 2 * self.out_channels, T, device=x.device, dtype=torch.float32)
        if B and T:
            with torch.cuda.device(x.device):
                ws = _lib.workspace(x.device, pack.workspace_bytes(B, T, prec))
                rc = _lib.lib().vsg_posterior_forward(pack.handle, xc.data_ptr(), mc.data_ptr(),
                                                      gc.data_ptr() if gc is not None else None, nc.data_ptr(), z.data_ptr(),
                                                      stats.data_ptr(), B, T, prec, ws.data_ptr(), ws.numel(),
                                                      _lib.stream_ptr(x.device))
            _lib.check(rc, "vsg_posterior_forward")
        mu_q, logs_q = torch.split(stats, self.out_channels, dim=1)
        return z, mu_q, logs_q

    def remove_weight_norm(self):
        self.enc.remove_weight_norm()
        self._vsg_pack = None


# ---- predictors (reference: modules/visinger/predictor.py) ------------------------------------------------------------

class PitchPredictor(nn.Module):
    """Frame-level (log-f0, voiced/unvoiced) head: a relative-position encoder conditioned on the speaker embedding,
    followed by a 1x1 projection (reference predictor.py:7-19).  Returns [B, T, out_dim]."""

    def __init__(self, in_dim, filter_channels, n_heads, n_layers, kernel_size, p_dropout, gin_channels, out_dim=2):
        super().__init__()
        self.pitch_predictor = RelativeEncoder(in_dim, filter_channels, n_heads, n_layers=n_layers, gin_channels=gin_channels,
                                               kernel_size=kernel_size, p_dropout=p_dropout)
        self.linear = nn.Conv1d(in_dim, out_dim, 1)

    def forward(self, x, x_mask, spk_emb):
        hidden = self.pitch_predictor(x, x_mask, g=spk_emb)
        return self.linear(hidden).transpose(1, 2)


class PhonemePredictor(nn.Module):
    """Parameter container for the training-only CTC head (reference predictor.py:22-35): present so that reference
    checkpoints load with strict key matching; it has no inference role."""

    def __init__(self, dict_size, hidden_channels, filter_channels, n_heads, n_layers, kernel_size, p_dropout):
        super().__init__()
        self.phoneme_predictor = RelativeEncoder(hidden_channels, filter_channels, n_heads, n_layers=n_layers,
                                                 kernel_size=kernel_size, p_dropout=p_dropout)
        self.ph_proj = nn.Conv1d(hidden_channels, dict_size, 1)

    def forward(self, *args, **kwargs):
        raise NotImplementedError("PhonemePredictor is training-only; visinger_b200 implements the inference path")

