"""CPU: the oracle's restatements of `MelSpectrogramFixed` (utils/audio/mel_processing.py:28-38) and of `save_wav`'s
arithmetic (utils/audio/io.py:8-14) against vectors the reference itself produced (tests/golden/make_golden_audio.py)."""
import numpy as np
import torch

from oracle import visinger_oracle as O
from helpers import load_npz


def test_mel_spectrogram_fixed_matches_reference():
    z = load_npz("audio_stage")
    wav = torch.from_numpy(z["wav"])
    mel = O.mel_spectrogram_fixed(wav)
    assert mel.shape == (wav.shape[0], 128, wav.shape[1] // 300)
    st = int(z["mel_stride"])
    assert float((mel[:, :, ::st] - torch.from_numpy(z["mel"])).abs().max()) <= 2e-5   # bit-exact on the same torch build
    assert O.mel_l1(wav, wav) == 0.0 and O.mel_l1(wav, 0.5 * wav) > 0.1


def test_mel_matches_torchaudio_if_present():
    try:
        from torchaudio.transforms import MelSpectrogram
    except Exception:
        import pytest
        pytest.skip("torchaudio not importable")
    w = torch.tanh(torch.randn(2, 6000, generator=torch.Generator().manual_seed(3)))
    m = MelSpectrogram(window_fn=torch.hann_window, **O.MEL_KW)
    ref = torch.log(m(w) + 0.001)[..., :-1]
    assert float((O.mel_spectrogram_fixed(w) - ref).abs().max()) <= 1e-6


def test_wav_to_int16_matches_reference_save_wav():
    z = load_npz("audio_stage")
    for b in range(z["wav"].shape[0]):
        got, peak = O.wav_to_int16(z["wav"][b], norm=True)
        assert got.dtype == np.int16 and np.array_equal(got, z["pcm"][b])
        assert peak == float(np.abs(z["wav"][b]).max())
        assert int(np.abs(got.astype(np.int32)).max()) == 32767          # the peak sample maps to full scale
    raw, _ = O.wav_to_int16(z["wav"][0], norm=False)
    assert np.array_equal(raw, (z["wav"][0] * 32767).astype(np.int16))
