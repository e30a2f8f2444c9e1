"""CPU oracle for the VISinger inference hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch *restatement* of the reference's algorithm for

    prior sampling -> ResidualCouplingBlock (reverse / forward) -> HiFi-GAN Generator

written as pure functions over a state-dict (name -> tensor).  It exists so that
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` have something to check / time the CUDA path against.  Nothing
under `visinger_b200/` may import it; the product path has no CPU fallback.

Arithmetic provenance: the reference's arithmetic lives in an un-vendored
third-party dependency, PyTorch (reference README pins torch==1.11.0+cu113; this
image has torch 2.11.0).  The primitives used below (`F.conv1d`,
`F.conv_transpose1d`, tanh, sigmoid, leaky_relu) are the very ATen CPU kernels the
reference's own `nn.Conv1d` / `nn.ConvTranspose1d` modules dispatch to, so the
fp32 oracle is bit-comparable with the reference run on CPU.  Weight-norm
(`w = g * v / ||v||`, recomputed on every forward by the reference's
`torch.nn.utils.weight_norm` pre-hook) is restated explicitly in `_eff_weight`.

Parity pinning: the reference ships NO tests, golden vectors or fixtures
(SURVEY.md section 4).  The oracle is therefore pinned against outputs of the
reference modules themselves, imported unmodified from /root/reference in the
build container by `tests/golden/make_golden.py`; the resulting fixtures live in
`tests/golden/*.npz` and `tests/test_oracle_golden.py` checks this file against
them (bit-exact in fp32 on the same torch build, <=1e-6 otherwise).

Reference citations are `path:line` relative to the reference repository root.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import math

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1  # modules/visinger/decoder.py:10

StateDict = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------
# weight handling
# --------------------------------------------------------------------------------------
def _eff_weight(sd: StateDict, prefix: str) -> torch.Tensor:
    """Effective conv weight for `prefix` ("...conv_name").

    Old-style weight norm (torch.nn.utils.weight_norm, dim=0) stores `weight_g`
    [D0,1,1] and `weight_v` [D0,D1,k] and recomputes w = v * (g / ||v||_{dims 1,2})
    before every forward (applied at modules/visinger/encoder.py:147,154,164 and
    modules/visinger/decoder.py:24-26,72-87,117-120).  For ConvTranspose1d dim 0 is
    C_in, so the norm runs over (C_out, k) -- SURVEY.md 7.2-7.  After
    `remove_weight_norm` only `.weight` exists; both spellings are accepted.
    """
    if prefix + ".weight" in sd:
        return sd[prefix + ".weight"]
    g = sd[prefix + ".weight_g"]
    v = sd[prefix + ".weight_v"]
    # torch._weight_norm(v, g, 0) == v * (g / ||v||_{dims 1,2}); it is the ATen primitive the
    # reference's pre-hook calls, used here so the fp32 oracle reproduces its rounding exactly.
    return torch._weight_norm(v, g, 0)


def _bias(sd: StateDict, prefix: str) -> Optional[torch.Tensor]:
    return sd.get(prefix + ".bias")


def get_padding(kernel_size: int, dilation: int = 1) -> int:
    """modules/commons/utils.py:109-110."""
    return int((kernel_size * dilation - dilation) / 2)


# --------------------------------------------------------------------------------------
# a1: prior sampling
# --------------------------------------------------------------------------------------
def prior_sample(mu_p, logs_p, noise, mask):
    """models/visinger.py:107 : z_p = (mu_p + randn_like(mu_p) * exp(logs_p)) * mask.

    The noise tensor is injected: CPU and CUDA generators differ.
    """
    return (mu_p + noise * torch.exp(logs_p)) * mask


# --------------------------------------------------------------------------------------
# a5/a6: WaveNet with the fused gate
# --------------------------------------------------------------------------------------
def gate(a, b, n_channels: int):
    """modules/visinger/encoder.py:206-213 fused_add_tanh_sigmoid_multiply."""
    s = a + b
    return torch.tanh(s[:, :n_channels]) * torch.sigmoid(s[:, n_channels:])


def wavenet(sd: StateDict, prefix: str, x, x_mask, g, *, hidden: int, kernel_size: int,
            dilation_rate: int, n_layers: int):
    """modules/visinger/encoder.py:167-195 WaveNet.forward (eval mode: dropout is identity)."""
    out = torch.zeros_like(x)
    if g is not None:
        g = F.conv1d(g, _eff_weight(sd, prefix + "cond_layer"), _bias(sd, prefix + "cond_layer"))
    for i in range(n_layers):
        dil = dilation_rate ** i
        pad = int((kernel_size * dil - dil) / 2)
        x_in = F.conv1d(x, _eff_weight(sd, f"{prefix}in_layers.{i}"), _bias(sd, f"{prefix}in_layers.{i}"),
                        dilation=dil, padding=pad)
        if g is not None:
            g_l = g[:, i * 2 * hidden:(i + 1) * 2 * hidden, :]
        else:
            g_l = torch.zeros_like(x_in)
        acts = gate(x_in, g_l, hidden)
        rs = F.conv1d(acts, _eff_weight(sd, f"{prefix}res_skip_layers.{i}"),
                      _bias(sd, f"{prefix}res_skip_layers.{i}"))
        if i < n_layers - 1:
            x = (x + rs[:, :hidden]) * x_mask
            out = out + rs[:, hidden:]
        else:
            out = out + rs
    return out * x_mask


# --------------------------------------------------------------------------------------
# f1: relative-position transformer encoder
# --------------------------------------------------------------------------------------
def channel_layer_norm(x, gamma, beta, eps: float = 1e-4):
    """modules/rel_transformer.py:33-42 LayerNorm.forward over dim 1 of [B, C, T]."""
    mean = torch.mean(x, 1, keepdim=True)
    variance = torch.mean((x - mean) ** 2, 1, keepdim=True)
    x = (x - mean) * torch.rsqrt(variance + eps)
    return x * gamma.view(1, -1, 1) + beta.view(1, -1, 1)


def rel_attention(sd: StateDict, prefix: str, x, attn_mask, *, n_heads: int, window: int):
    """modules/rel_transformer.py:123-177 MultiHeadAttention.forward (self-attention, heads_share=True, eval mode).  The
    reference moves the relative-position logits / weights between relative and absolute indexing with a pad-and-
    reshape trick (:218-243); stated directly: s_ij += q_i . Ek[j-i+w] and o_i += p_ij Ev[j-i+w] for |j-i| <= w."""
    q = F.conv1d(x, sd[prefix + "conv_q.weight"], sd[prefix + "conv_q.bias"])
    k = F.conv1d(x, sd[prefix + "conv_k.weight"], sd[prefix + "conv_k.bias"])
    v = F.conv1d(x, sd[prefix + "conv_v.weight"], sd[prefix + "conv_v.bias"])
    b, d, t = q.shape
    dk = d // n_heads
    q = q.view(b, n_heads, dk, t).transpose(2, 3)
    k = k.view(b, n_heads, dk, t).transpose(2, 3)
    v = v.view(b, n_heads, dk, t).transpose(2, 3)
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(dk)
    ek, ev = sd[prefix + "emb_rel_k"][0], sd[prefix + "emb_rel_v"][0]           # [2w+1, dk]
    idx = torch.arange(t)
    rel = idx[None, :] - idx[:, None] + window                                   # [t, t]: j - i + w
    inside = (rel >= 0) & (rel <= 2 * window)
    relc = rel.clamp(0, 2 * window)
    qe = torch.matmul(q, ek.t())                                                 # [b, h, t, 2w+1]
    local = torch.gather(qe, 3, relc.expand(b, n_heads, t, t)) * inside
    scores = scores + local / math.sqrt(dk)
    scores = scores.masked_fill(attn_mask == 0, -1e4)
    p = F.softmax(scores, dim=-1)
    out = torch.matmul(p, v)
    pw = torch.zeros(b, n_heads, t, 2 * window + 1, dtype=p.dtype)
    pw.scatter_add_(3, relc.expand(b, n_heads, t, t), p * inside)
    out = out + torch.matmul(pw, ev)
    out = out.transpose(2, 3).contiguous().view(b, d, t)
    return F.conv1d(out, sd[prefix + "conv_o.weight"], sd[prefix + "conv_o.bias"])


def rel_encoder(sd: StateDict, x, x_mask, g=None, *, n_heads: int = 2, n_layers: int = 4, kernel_size: int = 9,
                window: int = 4, prefix: str = ""):
    """modules/rel_transformer.py:286-320 RelativeEncoder.forward (pre_ln=False) with FFN :337-345 (ReLU)."""
    attn_mask = x_mask.unsqueeze(2) * x_mask.unsqueeze(-1)
    if g is not None:
        g = F.conv1d(g, sd[prefix + "pre_net.weight"], sd[prefix + "pre_net.bias"])
    for i in range(n_layers):
        if g is not None:
            x = x + g
        x = x * x_mask
        y = rel_attention(sd, f"{prefix}attn_layers.{i}.", x, attn_mask, n_heads=n_heads, window=window)
        x = channel_layer_norm(x + y, sd[f"{prefix}norm_layers_1.{i}.gamma"], sd[f"{prefix}norm_layers_1.{i}.beta"])
        h = F.conv1d(x * x_mask, sd[f"{prefix}ffn_layers.{i}.conv_1.weight"], sd[f"{prefix}ffn_layers.{i}.conv_1.bias"],
                     padding=kernel_size // 2)
        h = torch.relu(h)
        y = F.conv1d(h * x_mask, sd[f"{prefix}ffn_layers.{i}.conv_2.weight"], sd[f"{prefix}ffn_layers.{i}.conv_2.bias"])
        x = channel_layer_norm(x + y, sd[f"{prefix}norm_layers_2.{i}.gamma"], sd[f"{prefix}norm_layers_2.{i}.beta"])
    return x * x_mask


def rel_encoder_param_shapes(hidden=192, filter_channels=768, n_heads=2, n_layers=4, kernel_size=9, window=4, gin=None):
    """State-dict layout of RelativeEncoder (modules/rel_transformer.py:272-284)."""
    dk = hidden // n_heads
    shapes = {}
    for i in range(n_layers):
        a = f"attn_layers.{i}."
        for nm in ("conv_q", "conv_k", "conv_v", "conv_o"):
            shapes[a + nm + ".weight"] = (hidden, hidden, 1)
            shapes[a + nm + ".bias"] = (hidden,)
        shapes[a + "emb_rel_k"] = (1, 2 * window + 1, dk)
        shapes[a + "emb_rel_v"] = (1, 2 * window + 1, dk)
        shapes[f"ffn_layers.{i}.conv_1.weight"] = (filter_channels, hidden, kernel_size)
        shapes[f"ffn_layers.{i}.conv_1.bias"] = (filter_channels,)
        shapes[f"ffn_layers.{i}.conv_2.weight"] = (hidden, filter_channels, 1)
        shapes[f"ffn_layers.{i}.conv_2.bias"] = (hidden,)
        for j in (1, 2):
            shapes[f"norm_layers_{j}.{i}.gamma"] = (hidden,)
            shapes[f"norm_layers_{j}.{i}.beta"] = (hidden,)
    if gin:
        shapes["pre_net.weight"] = (hidden, gin, 1)
        shapes["pre_net.bias"] = (hidden,)
    return shapes


def synth_rel_encoder_state_dict(shapes: Dict[str, tuple], seed: int) -> StateDict:
    """Deterministic weights for the encoder: conv weights U(+-1/sqrt(fan_in)), biases U(+-0.05), relative embeddings
    N(0, 1/dk), LayerNorm gamma around 1 / beta around 0 (non-trivial, unlike the ones / zeros default init)."""
    gen = torch.Generator().manual_seed(seed)
    sd: StateDict = {}
    for name in sorted(shapes):
        shp = shapes[name]
        if name.endswith(".gamma"):
            sd[name] = 1.0 + 0.1 * (torch.rand(shp, generator=gen) * 2 - 1)
        elif name.endswith(".beta") or name.endswith(".bias"):
            sd[name] = (torch.rand(shp, generator=gen) * 2 - 1) * 0.05
        elif "emb_rel" in name:
            sd[name] = torch.randn(shp, generator=gen) * shp[-1] ** -0.5
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            sd[name] = (torch.rand(shp, generator=gen) * 2 - 1) * (1.0 / fan_in) ** 0.5
    return sd


# --------------------------------------------------------------------------------------
# f4: posterior encoder
# --------------------------------------------------------------------------------------
def posterior_encoder(sd: StateDict, x, x_mask, g, noise, *, out_channels: int = 192, hidden: int = 192,
                      kernel_size: int = 5, dilation_rate: int = 1, n_layers: int = 16, prefix: str = ""):
    """modules/visinger/encoder.py:92-98 PosteriorEncoder.forward with the noise injected (the reference draws
    torch.randn_like(mu_q) at :97).  Returns (z_q, mu_q, logs_q)."""
    h = F.conv1d(x, sd[prefix + "pre.weight"], sd[prefix + "pre.bias"]) * x_mask
    h = wavenet(sd, prefix + "enc.", h, x_mask, g, hidden=hidden, kernel_size=kernel_size, dilation_rate=dilation_rate,
                n_layers=n_layers)
    stats = F.conv1d(h, sd[prefix + "proj.weight"], sd[prefix + "proj.bias"]) * x_mask
    mu_q, logs_q = torch.split(stats, out_channels, dim=1)
    z_q = (mu_q + noise * torch.exp(logs_q)) * x_mask
    return z_q, mu_q, logs_q


# --------------------------------------------------------------------------------------
# a2/a3/a4: residual coupling block
# --------------------------------------------------------------------------------------
def coupling_layer(sd: StateDict, prefix: str, x, x_mask, g, reverse: bool, *, channels: int, hidden: int,
                   kernel_size: int, dilation_rate: int, n_layers: int):
    """modules/visinger/flow.py:66-85 ResidualCouplingLayer.forward with mean_only=True
    (the only way ResidualCouplingBlock builds it, flow.py:29-30): logs == 0."""
    half = channels // 2
    x0, x1 = x[:, :half], x[:, half:]
    h = F.conv1d(x0, sd[prefix + "pre.weight"], sd[prefix + "pre.bias"]) * x_mask
    h = wavenet(sd, prefix + "enc.", h, x_mask, g, hidden=hidden, kernel_size=kernel_size,
                dilation_rate=dilation_rate, n_layers=n_layers)
    m = F.conv1d(h, sd[prefix + "post.weight"], sd[prefix + "post.bias"]) * x_mask
    if not reverse:
        x1 = m + x1 * x_mask          # flow.py:78 with exp(logs) == 1
    else:
        x1 = (x1 - m) * x_mask        # flow.py:83 with exp(-logs) == 1
    return torch.cat([x0, x1], 1)


def flow(sd: StateDict, x, x_mask, g=None, reverse: bool = False, *, channels: int = 192, hidden: int = 192,
         kernel_size: int = 5, dilation_rate: int = 1, n_layers: int = 4, n_flows: int = 4, prefix: str = ""):
    """modules/visinger/flow.py:33-40 ResidualCouplingBlock.forward.

    flows[2i] is a coupling layer, flows[2i+1] a Flip (torch.flip over channels,
    flow.py:88-95).  Forward applies them in order, reverse in reversed order; both
    return only the tensor (logdet discarded at flow.py:36).
    """
    kw = dict(channels=channels, hidden=hidden, kernel_size=kernel_size, dilation_rate=dilation_rate,
              n_layers=n_layers)
    if not reverse:
        for i in range(n_flows):
            x = coupling_layer(sd, f"{prefix}flows.{2 * i}.", x, x_mask, g, False, **kw)
            x = torch.flip(x, [1])
    else:
        for i in reversed(range(n_flows)):
            x = torch.flip(x, [1])
            x = coupling_layer(sd, f"{prefix}flows.{2 * i}.", x, x_mask, g, True, **kw)
    return x


# --------------------------------------------------------------------------------------
# a7/a8/a9: HiFi-GAN generator
# --------------------------------------------------------------------------------------
def resblock1(sd: StateDict, prefix: str, x, kernel_size: int, dilations: Sequence[int]):
    """modules/visinger/decoder.py:91-104 ResBlock1.forward with x_mask=None (the only way
    Generator.forward calls it, decoder.py:51-53)."""
    for i, d in enumerate(dilations):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, _eff_weight(sd, f"{prefix}convs1.{i}"), _bias(sd, f"{prefix}convs1.{i}"),
                      dilation=d, padding=get_padding(kernel_size, d))
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, _eff_weight(sd, f"{prefix}convs2.{i}"), _bias(sd, f"{prefix}convs2.{i}"),
                      dilation=1, padding=get_padding(kernel_size, 1))
        x = xt + x
    return x


def resblock2(sd: StateDict, prefix: str, x, kernel_size: int, dilations: Sequence[int]):
    """modules/visinger/decoder.py:124-133 ResBlock2.forward with x_mask=None."""
    for i, d in enumerate(dilations):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, _eff_weight(sd, f"{prefix}convs.{i}"), _bias(sd, f"{prefix}convs.{i}"),
                      dilation=d, padding=get_padding(kernel_size, d))
        x = xt + x
    return x


def generator(sd: StateDict, x, g=None, *, resblock: str = "1", resblock_kernel_sizes=(3, 7, 11),
              resblock_dilation_sizes=((1, 3, 5), (1, 3, 5), (1, 3, 5)), upsample_rates=(5, 5, 3, 2, 2),
              upsample_kernel_sizes=(11, 11, 7, 4, 4), prefix: str = ""):
    """modules/visinger/decoder.py:40-59 Generator.forward."""
    nk = len(resblock_kernel_sizes)
    x = F.conv1d(x, sd[prefix + "conv_pre.weight"], sd[prefix + "conv_pre.bias"], padding=3)
    if g is not None:
        x = x + F.conv1d(g, sd[prefix + "cond.weight"], sd[prefix + "cond.bias"])
    rb = resblock1 if resblock == "1" else resblock2
    for i, (u, k) in enumerate(zip(upsample_rates, upsample_kernel_sizes)):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, _eff_weight(sd, f"{prefix}ups.{i}"), _bias(sd, f"{prefix}ups.{i}"),
                               stride=u, padding=(k - u) // 2)
        xs = None
        for j in range(nk):
            r = rb(sd, f"{prefix}resblocks.{i * nk + j}.", x, resblock_kernel_sizes[j],
                   resblock_dilation_sizes[j])
            xs = r if xs is None else xs + r
        x = xs / nk
    x = F.leaky_relu(x, LRELU_SLOPE)          # slope 0.1 here too (decoder.py:55)
    x = F.conv1d(x, sd[prefix + "conv_post.weight"], None, padding=3)
    return torch.tanh(x)


# --------------------------------------------------------------------------------------
# the whole hot path (models/visinger.py:105-111)
# --------------------------------------------------------------------------------------
def infer_hot_path(sd: StateDict, mu_p, logs_p, noise, mask, g, *, flow_kw=None, dec_kw=None,
                   flow_prefix: str = "flow.", dec_prefix: str = "decoder."):
    """models/visinger.py:107-111: sample, flow reverse, decode.  Returns (wav [B, L], z_q)."""
    z_p = prior_sample(mu_p, logs_p, noise, mask)
    z_q = flow(sd, z_p, mask, g=g, reverse=True, prefix=flow_prefix, **(flow_kw or {})) * mask
    wav = generator(sd, z_q * mask, g=g, prefix=dec_prefix, **(dec_kw or {})).squeeze(1)
    return wav, z_q


# --------------------------------------------------------------------------------------
# bf16-mode report metric: the reference's own log-mel spectrogram
# --------------------------------------------------------------------------------------
MEL_KW = dict(sample_rate=24000, n_fft=2048, win_length=1200, hop_length=300, f_min=20.0, f_max=12000.0, n_mels=128)
"""Parameters of `MelSpectrogramFixed` as the reference task builds it (tasks/visinger.py:32-35 with
config/datasets/svs/csd/preprocess.yaml:6-15: sample_rate 24000, fft_size 2048, win_size 1200, hop_size 300,
fmin 20, fmax 12000, num_mel_bins 128)."""


def _hz_to_mel_htk(f: float) -> float:
    import math
    return 2595.0 * math.log10(1.0 + f / 700.0)


def mel_filterbank(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale="htk") -- the filterbank behind
    torchaudio.transforms.MelSpectrogram's defaults, which utils/audio/mel_processing.py:33 instantiates."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_pts = torch.linspace(_hz_to_mel_htk(f_min), _hz_to_mel_htk(f_max), n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.min(down, up), min=0.0)            # [n_freqs, n_mels]


def mel_spectrogram_fixed(wav: torch.Tensor, *, sample_rate=24000, n_fft=2048, win_length=1200, hop_length=300,
                          f_min=20.0, f_max=12000.0, n_mels=128) -> torch.Tensor:
    """`MelSpectrogramFixed.forward` (utils/audio/mel_processing.py:28-38):
    log(MelSpectrogram(...)(x) + 0.001)[..., :-1], with torchaudio's defaults spelled out -- power spectrogram of a
    centred, reflect-padded STFT with a periodic Hann window of win_length zero-padded to n_fft, HTK mel filterbank
    without normalisation.  wav [..., L] -> [..., n_mels, L // hop_length]."""
    window = torch.hann_window(win_length, periodic=True, dtype=wav.dtype, device=wav.device)
    shape = wav.shape
    x = wav.reshape(-1, shape[-1])
    spec = torch.stft(x, n_fft, hop_length=hop_length, win_length=win_length, window=window, center=True,
                      pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
    power = spec.abs().pow(2.0)                                 # [N, n_freqs, frames]
    fb = mel_filterbank(n_fft // 2 + 1, f_min, f_max, n_mels, sample_rate).to(wav.dtype).to(wav.device)
    mel = torch.matmul(power.transpose(-1, -2), fb).transpose(-1, -2)
    out = torch.log(mel + 0.001)[..., :-1]
    return out.reshape(shape[:-1] + out.shape[-2:])


def mel_l1(wav: torch.Tensor, wav_ref: torch.Tensor) -> float:
    """Mean absolute difference of the two log-mel spectrograms: the reference's own mel loss without its weight
    (tasks/visinger.py: F.l1_loss(mel_fn(wav_out), mel)); the bf16-mode report metric north_star names."""
    return float((mel_spectrogram_fixed(wav.float()) - mel_spectrogram_fixed(wav_ref.float())).abs().mean())


# --------------------------------------------------------------------------------------
# output stage: utils/audio/io.py:8-14 up to the file write
# --------------------------------------------------------------------------------------
def wav_to_int16(wav, norm: bool = True):
    """`save_wav` (utils/audio/io.py:8-14) without the file write: `wav / np.abs(wav).max()` when norm
    (out_wav_norm: true, config/models/base_task.yaml:53), `* 32767`, `.astype(np.int16)` -- float32 numpy
    arithmetic, conversion truncating toward zero.  wav: 1-D float32 array (ONE utterance, valid samples only:
    the reference runs batch 1, tasks/visinger.py:246).  Returns (int16 array, peak)."""
    import numpy as np
    w = np.asarray(wav, dtype=np.float32)
    peak = np.abs(w).max() if w.size else np.float32(0)
    if norm:
        w = w / peak
    w = w * 32767
    return w.astype(np.int16), float(peak)


# --------------------------------------------------------------------------------------
# deterministic synthetic weights (used by tests, smoke and bench on both sides)
# --------------------------------------------------------------------------------------
def flow_param_shapes(channels=192, hidden=192, kernel_size=5, n_layers=4, n_flows=4, gin=256):
    """State-dict layout of ResidualCouplingBlock."""
    shapes = {}
    half = channels // 2
    for f in range(n_flows):
        p = f"flows.{2 * f}."
        shapes[p + "pre.weight"] = (hidden, half, 1)
        shapes[p + "pre.bias"] = (hidden,)
        if gin:
            shapes[p + "enc.cond_layer.bias"] = (2 * hidden * n_layers,)
            shapes[p + "enc.cond_layer.weight_g"] = (2 * hidden * n_layers, 1, 1)
            shapes[p + "enc.cond_layer.weight_v"] = (2 * hidden * n_layers, gin, 1)
        for i in range(n_layers):
            shapes[p + f"enc.in_layers.{i}.bias"] = (2 * hidden,)
            shapes[p + f"enc.in_layers.{i}.weight_g"] = (2 * hidden, 1, 1)
            shapes[p + f"enc.in_layers.{i}.weight_v"] = (2 * hidden, hidden, kernel_size)
            rs = 2 * hidden if i < n_layers - 1 else hidden
            shapes[p + f"enc.res_skip_layers.{i}.bias"] = (rs,)
            shapes[p + f"enc.res_skip_layers.{i}.weight_g"] = (rs, 1, 1)
            shapes[p + f"enc.res_skip_layers.{i}.weight_v"] = (rs, hidden, 1)
        shapes[p + "post.weight"] = (half, hidden, 1)
        shapes[p + "post.bias"] = (half,)
    return shapes


def posterior_param_shapes(in_channels=1025, out_channels=192, hidden=192, kernel_size=5, n_layers=16, gin=256):
    """State-dict layout of PosteriorEncoder (modules/visinger/encoder.py:77-90)."""
    shapes = {"pre.weight": (hidden, in_channels, 1), "pre.bias": (hidden,),
              "proj.weight": (2 * out_channels, hidden, 1), "proj.bias": (2 * out_channels,)}
    if gin:
        shapes["enc.cond_layer.bias"] = (2 * hidden * n_layers,)
        shapes["enc.cond_layer.weight_g"] = (2 * hidden * n_layers, 1, 1)
        shapes["enc.cond_layer.weight_v"] = (2 * hidden * n_layers, gin, 1)
    for i in range(n_layers):
        shapes[f"enc.in_layers.{i}.bias"] = (2 * hidden,)
        shapes[f"enc.in_layers.{i}.weight_g"] = (2 * hidden, 1, 1)
        shapes[f"enc.in_layers.{i}.weight_v"] = (2 * hidden, hidden, kernel_size)
        rs = 2 * hidden if i < n_layers - 1 else hidden
        shapes[f"enc.res_skip_layers.{i}.bias"] = (rs,)
        shapes[f"enc.res_skip_layers.{i}.weight_g"] = (rs, 1, 1)
        shapes[f"enc.res_skip_layers.{i}.weight_v"] = (rs, hidden, 1)
    return shapes


def generator_param_shapes(initial_channel=192, resblock="1", resblock_kernel_sizes=(3, 7, 11),
                           resblock_dilation_sizes=((1, 3, 5),) * 3, upsample_rates=(5, 5, 3, 2, 2),
                           upsample_initial_channel=512, upsample_kernel_sizes=(11, 11, 7, 4, 4), gin=256):
    """State-dict layout of Generator."""
    shapes = {"conv_pre.weight": (upsample_initial_channel, initial_channel, 7),
              "conv_pre.bias": (upsample_initial_channel,)}
    ch = upsample_initial_channel
    nk = len(resblock_kernel_sizes)
    for i, (u, k) in enumerate(zip(upsample_rates, upsample_kernel_sizes)):
        cin, ch = upsample_initial_channel // 2 ** i, upsample_initial_channel // 2 ** (i + 1)
        shapes[f"ups.{i}.bias"] = (ch,)
        shapes[f"ups.{i}.weight_g"] = (cin, 1, 1)
        shapes[f"ups.{i}.weight_v"] = (cin, ch, k)
        for j, (rk, rd) in enumerate(zip(resblock_kernel_sizes, resblock_dilation_sizes)):
            p = f"resblocks.{i * nk + j}."
            names = ("convs1", "convs2") if resblock == "1" else ("convs",)
            for nm in names:
                for q in range(len(rd)):
                    shapes[p + f"{nm}.{q}.bias"] = (ch,)
                    shapes[p + f"{nm}.{q}.weight_g"] = (ch, 1, 1)
                    shapes[p + f"{nm}.{q}.weight_v"] = (ch, ch, rk)
    shapes["conv_post.weight"] = (1, ch, 7)
    if gin:
        shapes["cond.weight"] = (upsample_initial_channel, gin, 1)
        shapes["cond.bias"] = (upsample_initial_channel,)
    return shapes


def synth_state_dict(shapes: Dict[str, tuple], seed: int) -> StateDict:
    """Deterministic random weights keyed by name.

    Mimics the reference's effective default init (kaiming-uniform v, g ~ ||v||,
    small uniform bias) and -- unlike the reference init
    (flow.py:63-64 zeroes `post`) -- makes every coupling layer non-trivial
    (SURVEY.md Appendix B-2).  Tensors are drawn in
    sorted-name order from one torch.Generator so both sides of a parity test can
    rebuild identical weights from the seed alone.
    """
    gen = torch.Generator().manual_seed(seed)
    sd: StateDict = {}
    for name in sorted(shapes):
        shp = shapes[name]
        if name.endswith("weight_g"):
            continue  # derived from v below
        if name.endswith(".bias"):
            sd[name] = (torch.rand(shp, generator=gen) * 2 - 1) * 0.05
        elif ".post." in name:
            # SURVEY.md 8(c) patch (2): re-randomise the zero-initialised `post` at std 0.05
            sd[name] = (torch.rand(shp, generator=gen) * 2 - 1) * (0.05 * 3 ** 0.5)
        else:
            # PyTorch's default conv init: kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)),
            # fan_in = shape[1] * k (which for ConvTranspose1d is C_out * k, as torch computes it)
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            bound = (1.0 / fan_in) ** 0.5
            sd[name] = (torch.rand(shp, generator=gen) * 2 - 1) * bound
    for name in sorted(shapes):
        if name.endswith("weight_g"):
            v = sd[name[:-1] + "v"]
            norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(shapes[name])
            # g = ||v|| * (1 +- 10 %): exercises the weight-norm fold without changing the scale
            sd[name] = norm * (0.9 + 0.2 * torch.rand(shapes[name], generator=gen))
    return sd
