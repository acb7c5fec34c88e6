"""log-mel + MFCC front-end, fp64 CPU statement.  TEST INFRASTRUCTURE (see oracle/__init__.py).

**PARITY UNPINNED against the reference**: the reference contains no STFT / mel /
MFCC code at all (src/dataprocessing.py:184-212 only loads, pads
and peak-normalises the waveform; "MFCC" occurs once, in the docstring at
src/encoders.py:335).  The parameters BASELINE.json fixes are sr=16000,
n_fft=512, hop=160, 64 mels, 40 MFCC, 3 s clips.  Everything else below is this
repo's written spec, chosen to follow the conventions of the reference's declared
dependency librosa==0.10.1 (pyproject.toml:35; not vendored, not installed here)
as published in its documentation: librosa.stft(center=True, pad_mode="reflect"
... here applied after pre-emphasis), filters.mel(htk=False, norm="slaney"),
power_to_db(ref=1, amin=1e-10, top_db=None), scipy dct(type=2, norm="ortho").
It is cross-checked in tests/test_oracle_frontend.py against the independent
implementations available in this image (torch.stft,
torchaudio.functional.melscale_fbanks, scipy.fft.dct).

Spec, per clip y[0..N-1] (N = 48000):
  1. pre-emphasis      p[n] = y[n] - 0.97 * y[n-1],  y[-1] := 0
  2. centred framing   q = reflect-pad(p, 256, 256)  (numpy "reflect": no edge repeat)
                       frame t = q[160 t : 160 t + 512],  t = 0 .. N // 160   (301 frames)
  3. window            periodic Hann  w[n] = 0.5 - 0.5 cos(2 pi n / 512)
  4. power spectrum    S[k] = |rfft(frame * w)[k]|^2,  k = 0..256
  5. mel               M = melfb(64, 257) @ S      Slaney scale, Slaney area norm, 0..8000 Hz
  6. log               L = 10 log10(max(M, 1e-10))
  7. MFCC              C = dct_ortho_II(L)[:40]
  8. output row t      [L (64) | C (40)]  ->  (301, 104), row-major (feeds
                       SequenceEncoder batch_first, src/encoders.py:58,127)
"""
from __future__ import annotations

import numpy as np

SR = 16000
N_FFT = 512
HOP = 160
N_MELS = 64
N_MFCC = 40
N_BINS = N_FFT // 2 + 1
PREEMPH = 0.97
LOG_FLOOR = 1e-10
F_MIN = 0.0
F_MAX = 8000.0
N_FEATS = N_MELS + N_MFCC


def n_frames(n_samples: int) -> int:
    return 1 + n_samples // HOP


def hz_to_mel_slaney(f):
    """Slaney (Auditory Toolbox) mel scale: linear below 1 kHz, log above."""
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz,
                    min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, mels)


def mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(n_mels=N_MELS, n_fft=N_FFT, sr=SR, fmin=F_MIN, fmax=F_MAX) -> np.ndarray:
    """(n_mels, n_fft//2+1) triangular filters, Slaney area normalisation (fp64)."""
    fftfreqs = np.linspace(0.0, sr / 2.0, n_fft // 2 + 1)
    mel_pts = np.linspace(hz_to_mel_slaney(fmin), hz_to_mel_slaney(fmax), n_mels + 2)
    hz_pts = mel_to_hz_slaney(mel_pts)
    fdiff = np.diff(hz_pts)
    ramps = hz_pts[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    fb = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (hz_pts[2:n_mels + 2] - hz_pts[:n_mels])
    return fb * enorm[:, None]


def dct_matrix(n_out=N_MFCC, n_in=N_MELS) -> np.ndarray:
    """Orthonormal DCT-II rows: C[c, j] = s_c cos(pi (j + 1/2) c / n_in)."""
    j = np.arange(n_in)[None, :]
    c = np.arange(n_out)[:, None]
    m = np.cos(np.pi * (j + 0.5) * c / n_in) * np.sqrt(2.0 / n_in)
    m[0] *= np.sqrt(0.5)
    return m


def hann_periodic(n=N_FFT) -> np.ndarray:
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def preemphasis(y: np.ndarray) -> np.ndarray:
    y = np.asarray(y, dtype=np.float64)
    p = y.copy()
    p[..., 1:] -= PREEMPH * y[..., :-1]
    return p


def power_spectrogram(wave: np.ndarray) -> np.ndarray:
    """(B, N) -> (B, T, 257) fp64 power spectrum of the pre-emphasised, framed, windowed signal."""
    p = preemphasis(wave)
    q = np.pad(p, ((0, 0), (N_FFT // 2, N_FFT // 2)), mode="reflect")
    T = n_frames(wave.shape[-1])
    idx = (np.arange(T) * HOP)[:, None] + np.arange(N_FFT)[None, :]
    frames = q[:, idx] * hann_periodic()[None, None, :]
    return np.abs(np.fft.rfft(frames, axis=-1)) ** 2


def logmel_mfcc(wave: np.ndarray, out_dtype=np.float32) -> np.ndarray:
    """(B, N) fp32/fp64 waveform -> (B, 1 + N//160, 104) features [64 log-mel | 40 MFCC]."""
    wave = np.atleast_2d(np.asarray(wave, dtype=np.float64))
    S = power_spectrogram(wave)
    M = S @ mel_filterbank().T
    L = 10.0 * np.log10(np.maximum(M, LOG_FLOOR))
    C = L @ dct_matrix().T
    return np.concatenate([L, C], axis=-1).astype(out_dtype)
