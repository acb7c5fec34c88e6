"""oracle/frontend.py has no reference to pin against (parity unpinned).  It is cross-checked here
against the independent building blocks in this image: torch.stft, torchaudio's Slaney filterbank, scipy's DCT,
and against the committed golden features (drift guard)."""
import numpy as np
import pytest
import torch

from oracle import frontend as fe, synth


def _wave(B=3, seed=7):
    return synth.synth_audio(np.arange(B, dtype=np.uint64) + 11, seed)


def test_shapes_and_frames():
    w = _wave(2)
    f = fe.logmel_mfcc(w)
    assert f.shape == (2, 301, 104) and f.dtype == np.float32
    assert fe.n_frames(48000) == 301
    assert fe.logmel_mfcc(w[:, :1600]).shape == (2, 11, 104)


def test_power_spectrum_matches_torch_stft():
    w = _wave(2)
    p = fe.preemphasis(w)
    S = torch.stft(torch.from_numpy(p), n_fft=512, hop_length=160, window=torch.hann_window(512, periodic=True,
                                                                                         dtype=torch.float64),
                   center=True, pad_mode="reflect", return_complex=True)
    want = (S.abs() ** 2).transpose(1, 2).numpy()
    got = fe.power_spectrogram(w)
    assert got.shape == want.shape == (2, 301, 257)
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-12 * want.max())


def test_mel_filterbank_matches_torchaudio_slaney():
    ta = pytest.importorskip("torchaudio")
    want = ta.functional.melscale_fbanks(257, 0.0, 8000.0, 64, 16000, norm="slaney", mel_scale="slaney").T.numpy()
    got = fe.mel_filterbank()
    assert got.shape == (64, 257)
    np.testing.assert_allclose(got, want, rtol=2e-4, atol=1e-7)   # torchaudio builds it in fp32
    assert (got >= 0).all() and (got.sum(axis=1) > 0).all()


def test_dct_matches_scipy():
    from scipy.fft import dct
    x = np.random.default_rng(0).normal(size=(5, 64))
    want = dct(x, type=2, norm="ortho", axis=-1)[:, :40]
    np.testing.assert_allclose(x @ fe.dct_matrix().T, want, rtol=1e-12, atol=1e-12)


def test_log_floor_on_silence():
    f = fe.logmel_mfcc(np.zeros((1, 48000), dtype=np.float32))
    assert np.allclose(f[0, :, :64], -100.0)                # 10 log10(1e-10)
    assert np.allclose(f[0, :, 64], -100.0 * 8.0, atol=1e-3)  # DC coefficient: sum / sqrt(64)
    assert np.allclose(f[0, :, 65:], 0.0, atol=1e-3)


def test_golden_features(golden):
    ids = golden["clip_ids"]
    w = synth.synth_audio(ids[[0, -1]], int(golden["seed"]))
    f = fe.logmel_mfcc(w)
    np.testing.assert_allclose(f[0], golden["feats_clip0"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(f[1], golden["feats_last"], rtol=0, atol=1e-4)


def test_whole_front_end_matches_torchaudio_pipeline():
    """End to end against an independent implementation of the same conventions: torchaudio's MelSpectrogram (Slaney
    scale and normalisation - the options torchaudio added to reproduce librosa's filters.mel), AmplitudeToDB (power,
    amin 1e-10, no top_db = librosa.power_to_db(ref=1, top_db=None)) and create_dct(norm="ortho") (= scipy dct type 2,
    what librosa.feature.mfcc applies), on the pre-emphasised waveform."""
    ta = pytest.importorskip("torchaudio")
    w = _wave(2)
    p = torch.from_numpy(fe.preemphasis(w))                                     # fp64
    mel = ta.transforms.MelSpectrogram(sample_rate=16000, n_fft=512, hop_length=160, f_min=0.0, f_max=8000.0, n_mels=64,
                                       window_fn=torch.hann_window, power=2.0, center=True, pad_mode="reflect",
                                       norm="slaney", mel_scale="slaney").double()
    logmel = ta.transforms.AmplitudeToDB(stype="power", top_db=None)(mel(p))    # (B, 64, 301)
    dct = ta.functional.create_dct(40, 64, norm="ortho").double()               # (64, 40)
    mfcc = logmel.transpose(1, 2) @ dct
    want = torch.cat([logmel.transpose(1, 2), mfcc], dim=-1).numpy()
    got = fe.logmel_mfcc(w, out_dtype=np.float64)
    assert got.shape == want.shape == (2, 301, 104)
    # torchaudio builds its filterbank in fp32 (2e-4 relative on a tap): a few 1e-3 dB on a log-mel value
    assert np.abs(got - want).max() <= 5e-3
    assert np.abs(got - want).max() / np.abs(want).max() <= 1e-5
