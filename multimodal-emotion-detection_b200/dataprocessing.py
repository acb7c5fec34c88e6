"""Audio side of src/dataprocessing.py on the device.

* ``condition_audio``  the post-conditions of ``load_raw_audio`` (src/dataprocessing.py:198-211): fixed
  length (truncate / right zero pad), divide by max|y| when > 0, fp32.
* ``load_raw_audio``   same signature as the reference (src/dataprocessing.py:184-188).  The reference
  decodes and resamples with ``librosa.load`` (:196); file decode is outside the hot path, so this reads
  PCM/float WAV files whose sample rate already equals ``sr`` and conditions them on the device.
* ``logmel_mfcc``      NEW - the reference has no front-end.  (B, N) waveform ->
  (B, 1 + N // 160, 104) = [64 log-mel | 40 MFCC]; spec in oracle/frontend.py.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

SR = 16000
N_FFT = 512
HOP = 160
N_MELS = 64
N_MFCC = 40
N_FEATS = N_MELS + N_MFCC


def n_frames(n_samples: int) -> int:
    return 1 + n_samples // HOP


def condition_audio(wave: torch.Tensor, lengths: torch.Tensor | None = None, n_out: int | None = None) -> torch.Tensor:
    """wave (B, N_in) CUDA -> (B, n_out) fp32: first min(len_b, n_out) samples, right zero padded, divided by
    max|.| when > 0 (src/dataprocessing.py:198-211).  int16 input is 16-bit PCM (sample k = k / 32768, what
    librosa.load makes of a 16-bit .wav, :196): same result, bit for bit, as converting first."""
    dev = _lib.require_cuda(wave, lengths)
    _lib.check_device(dev)
    if wave.dim() != 2:
        raise ValueError(f"wave must be (B, N), got {tuple(wave.shape)}")
    pcm = wave.dtype == torch.int16
    wave = wave.contiguous() if pcm else _lib.f32c(wave)
    B, n_in = wave.shape
    n_out = int(n_out if n_out is not None else n_in)
    if lengths is not None:
        lengths = lengths.to(device=dev, dtype=torch.int32).contiguous()
        if lengths.shape != (B,):
            raise ValueError("lengths must be (B,)")
    out = torch.empty(B, n_out, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        fn = _lib.load().avhot_condition_audio_i16 if pcm else _lib.load().avhot_condition_audio
        _lib.check(fn(_lib.ptr(wave), _lib.ptr(lengths), _lib.ptr(out), B, n_in, n_out, _lib.stream_ptr(dev)),
                   "condition_audio")
    return out


def load_raw_audio(wav_path: str, sr: int = 16000, max_duration: float = 3.0) -> np.ndarray:
    """Same contract as the reference: (T, 1) float32, T = sr * max_duration, peak-normalised."""
    from scipy.io import wavfile
    file_sr, y = wavfile.read(wav_path)
    if file_sr != sr:
        raise ValueError(f"{wav_path}: sample rate {file_sr} != {sr}; resampling (librosa.load in the reference, "
                         "src/dataprocessing.py:196) is outside this library's hot path")
    if y.ndim == 2:
        y = y.mean(axis=1)                                 # librosa.load(mono=True)
    if np.issubdtype(y.dtype, np.integer):
        y = y.astype(np.float32) / float(np.iinfo(y.dtype).max + 1)
    target = int(sr * max_duration)
    w = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32)).cuda()[None, :]
    out = condition_audio(w, n_out=target)
    return out[0].cpu().numpy().reshape(-1, 1)


def logmel_mfcc(wave: torch.Tensor) -> torch.Tensor:
    """(B, N) or (B, N, 1) fp32 CUDA waveform -> (B, 1 + N // 160, 104) features."""
    dev = _lib.require_cuda(wave)
    _lib.check_device(dev)
    if wave.dim() == 3 and wave.shape[-1] == 1:
        wave = wave[..., 0]
    if wave.dim() != 2:
        raise ValueError(f"wave must be (B, N) or (B, N, 1), got {tuple(wave.shape)}")
    wave = _lib.f32c(wave)
    B, N = wave.shape
    feats = torch.empty(B, n_frames(N), N_FEATS, device=dev, dtype=torch.float32)
    lib = _lib.load()
    with torch.cuda.device(dev):
        for b0 in range(0, max(B, 1), 32768):              # the kernel takes <= 65535 clips per launch
            nb = min(32768, B - b0)
            if nb <= 0:
                break
            _lib.check(lib.avhot_frontend(_lib.ptr(wave[b0:]), _lib.ptr(feats[b0:]), nb, N, _lib.stream_ptr(dev)),
                       "logmel_mfcc")
    return feats


def frontend_tables():
    """The fp32 tables the kernel uses: mel (64, 257), dct (40, 64), window (512,) as numpy arrays."""
    mel = np.empty((N_MELS, N_FFT // 2 + 1), dtype=np.float32)
    dct = np.empty((N_MFCC, N_MELS), dtype=np.float32)
    win = np.empty((N_FFT,), dtype=np.float32)
    _lib.check(_lib.load().avhot_frontend_tables_host(mel.ctypes.data_as(C.c_void_p), dct.ctypes.data_as(C.c_void_p),
                                                      win.ctypes.data_as(C.c_void_p)))
    return mel, dct, win
