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
        nonpad = (text_tokens > 0).float().unsqueeze(1)                      # [B, 1, T_ph]
        emb = torch.cat([self.ph_emb(text_tokens), self.pitch_emb(pitch_tokens), self.dur_emb(dur_tokens)], 2)
        tok = self.linear(emb * self.embed_scale) * nonpad.transpose(1, 2)   # [B, T_ph, H]
        if self.use_pos_embed:
            # Reference quirk kept bit-for-bit (encoder.py:50-53): seq_len is passed as tok.shape[2] (= H, not T_ph), so
            # the [B*T_ph, H] table lookup is *viewed* as [B, H, T_ph] and then transposed -- a scrambled embedding.
            pos = self.embed_positions(tok.shape[0], tok.shape[2], tok[..., 0])
            tok = tok + pos.transpose(1, 2)
        tok = tok * nonpad.transpose(1, 2)
        enc = self.text_encoder(tok.transpose(1, 2), nonpad)                 # [B, H, T_ph]
        return expand_states(enc.transpose(1, 2), mel2ph).transpose(1, 2)    # [B, H, T]


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


class PosteriorEncoder(nn.Module):
    """Training-only in the reference (models/visinger.py:94).  Parameter container so full checkpoints load."""

    def __init__(self, in_channels, out_channels, hidden_channels, kernel_size, dilation_rate, n_layers, gin_channels):
        super().__init__()
        self.out_channels = out_channels
        self.pre = nn.Conv1d(in_channels, hidden_channels, 1)
        self.enc = WaveNet(hidden_channels, kernel_size, dilation_rate, n_layers, gin_channels=gin_channels)
        self.proj = nn.Conv1d(hidden_channels, out_channels * 2, 1)

    def forward(self, *a, **k):
        raise NotImplementedError("PosteriorEncoder is training-only; visinger_b200 implements the inference path")


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

