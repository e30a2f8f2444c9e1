"""Output stage of the hot path: mirror of the reference's `utils/audio/io.py::save_wav` (lines 8-14).

The reference pulls the fp32 waveform to the host (`inference/visinger.py:98`), peak-normalises it with numpy, scales by
32767, casts to int16 and writes a WAV.  Here the normalise / scale / cast runs on the GPU (`vsg_wav_to_int16`,
csrc/output.cu), bit-identical to the numpy arithmetic, so only int16 PCM crosses PCIe.  There is no CPU fallback.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ... import _lib


@torch.no_grad()
def wav_to_int16(wav: torch.Tensor, lengths: Optional[torch.Tensor] = None, norm: bool = True,
                 out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """wav: CUDA fp32 [B, L] or [B, 1, L] (the decoder's output); lengths: valid SAMPLES per utterance ([B], any int
    dtype / device) or None.  Returns (pcm int16 [B, L] on the device, peak fp32 [B])."""
    _lib.require_cuda(wav, "wav")
    w = _lib.as_f32c(wav.reshape(wav.shape[0], -1))
    B, L = w.shape
    dev = w.device
    pcm = out if out is not None else torch.empty(B, L, dtype=torch.int16, device=dev)
    if pcm.dtype != torch.int16 or pcm.shape != (B, L) or not pcm.is_cuda or not pcm.is_contiguous():
        raise RuntimeError("out must be a contiguous CUDA int16 tensor of shape [B, L]")
    peak = torch.empty(B, dtype=torch.float32, device=dev)
    ln = None
    if lengths is not None:
        if lengths.numel() != B:
            raise RuntimeError(f"lengths must have {B} entries, got {lengths.numel()}")
        ln = lengths.to(device=dev, dtype=torch.int32).contiguous()
    with torch.cuda.device(dev):
        rc = _lib.lib().vsg_wav_to_int16(w.data_ptr(), ln.data_ptr() if ln is not None else None, pcm.data_ptr(),
                                         peak.data_ptr(), B, L, 1 if norm else 0, _lib.stream_ptr(dev))
    _lib.check(rc, "vsg_wav_to_int16")
    return pcm, peak


def save_wav(wav, path: str, sr: int, norm: bool = False) -> None:
    """Reference signature (`utils/audio/io.py:8`): ONE utterance, written as 16-bit PCM.  `wav` is a CUDA tensor
    [L] (the reference takes the numpy copy of the same data)."""
    from scipy.io import wavfile
    if not isinstance(wav, torch.Tensor):
        raise TypeError("visinger_b200 save_wav takes the CUDA waveform tensor (there is no CPU path)")
    pcm, _ = wav_to_int16(wav.reshape(1, -1), None, norm)
    wavfile.write(path[:-4] + ".wav", sr, pcm[0].cpu().numpy())
