"""Golden vectors for the two audio-side functions the hot path's reports / output stage follow, generated from the
UNMODIFIED reference (build container only; /root/reference does not exist on the GPU box):

    python tests/golden/make_golden_audio.py

  * `MelSpectrogramFixed` (utils/audio/mel_processing.py:28-38) with the task's parameters (tasks/visinger.py:32-35):
    the log-mel spectrogram behind the bf16-mode "mel-spectrogram L1" report;
  * `save_wav(norm=True)` (utils/audio/io.py:8-14): peak-normalise, x 32767, int16 -- written to a temporary WAV by
    the reference itself and read back.

Stores the seeded input waveform, the reference log-mel (every 3rd frame) and the int16 samples in `audio_stage.npz`.
"""
import os
import sys
import tempfile
import types
import warnings

import numpy as np
import torch
from scipy.io import wavfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("VISINGER_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

import utils  # noqa: E402  (reference package)
# utils/audio/__init__.py pulls librosa / webrtcvad (absent here): register the package without running it
pkg = types.ModuleType("utils.audio")
pkg.__path__ = [os.path.join(REF, "utils", "audio")]
sys.modules["utils.audio"] = pkg
from utils.audio.mel_processing import MelSpectrogramFixed  # noqa: E402  (reference)
from utils.audio.io import save_wav  # noqa: E402  (reference)

from oracle import visinger_oracle as O  # noqa: E402

if __name__ == "__main__":
    gen = torch.Generator().manual_seed(4321)
    B, L = 3, 9000
    t = torch.arange(L) / 24000.0
    wav = 0.6 * torch.sin(2 * torch.pi * 220.0 * t)[None] * torch.rand(B, 1, generator=gen) + 0.1 * torch.randn(B, L, generator=gen)
    wav = torch.tanh(wav)
    mel_fn = MelSpectrogramFixed(sample_rate=24000, n_fft=2048, win_length=1200, hop_length=300, f_min=20, f_max=12000,
                                 n_mels=128, window_fn=torch.hann_window)
    mel = mel_fn(wav)
    pcm = []
    with tempfile.TemporaryDirectory() as d:
        for b in range(B):
            path = os.path.join(d, f"u{b}.wav")
            save_wav(wav[b].numpy(), path, 24000, norm=True)
            sr, data = wavfile.read(path)
            assert sr == 24000 and data.dtype == np.int16
            pcm.append(data)
    pcm = np.stack(pcm)
    mine = O.mel_spectrogram_fixed(wav)
    print("oracle mel vs reference:", float((mine - mel).abs().max()))
    for b in range(B):
        got, _ = O.wav_to_int16(wav[b].numpy(), norm=True)
        print("oracle int16 vs reference:", int(np.abs(got.astype(np.int32) - pcm[b].astype(np.int32)).max()))
    np.savez_compressed(os.path.join(HERE, "audio_stage.npz"), wav=wav.numpy(), mel=mel[:, :, ::3].numpy(), mel_stride=3,
                        pcm=pcm)
