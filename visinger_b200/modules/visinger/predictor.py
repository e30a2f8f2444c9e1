"""PyTorch mirrors of the reference's `modules/visinger/predictor.py`: PitchPredictor (:7-19) and a parameter-only
PhonemePredictor (:22-35, training-only CTC head)."""
from __future__ import annotations

import torch.nn as nn

from ..rel_transformer import RelativeEncoder


class PitchPredictor(nn.Module):
    def __init__(self, in_dim, filter_channels, n_heads, n_layers, kernel_size, p_dropout, gin_channels, out_dim=2):
        super().__init__()
        self.pitch_predictor = RelativeEncoder(in_dim, filter_channels, n_heads, n_layers=n_layers, gin_channels=gin_channels,
                                               kernel_size=kernel_size, p_dropout=p_dropout)
        self.linear = nn.Conv1d(in_dim, out_dim, 1)

    def forward(self, x, x_mask, spk_emb):
        return self.linear(self.pitch_predictor(x, x_mask, g=spk_emb)).transpose(1, 2)   # [B, T, out_dim]


class PhonemePredictor(nn.Module):
    def __init__(self, dict_size, hidden_channels, filter_channels, n_heads, n_layers, kernel_size, p_dropout):
        super().__init__()
        self.phoneme_predictor = RelativeEncoder(hidden_channels, filter_channels, n_heads, n_layers=n_layers,
                                                 kernel_size=kernel_size, p_dropout=p_dropout)
        self.ph_proj = nn.Conv1d(hidden_channels, dict_size, 1)

    def forward(self, *a, **k):
        raise NotImplementedError("PhonemePredictor is training-only; visinger_b200 implements the inference path")
