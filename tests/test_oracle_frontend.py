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
