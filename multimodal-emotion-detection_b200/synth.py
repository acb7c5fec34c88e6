"""Seeded synthetic RAVDESS-shaped inputs generated on the device, bit-exact with oracle/synth.py (every value is
a pure function of (seed, global clip index, element), so shards agree at any world size)."""
from __future__ import annotations

import torch

from . import _lib

N_SAMPLES = 48000
N_FRAMES_VIDEO = 16
FRAME_DIM = 112 * 112


def synth_audio(B: int, clip0: int, seed: int, device, n_samples: int = N_SAMPLES) -> torch.Tensor:
    dev = torch.device(device)
    out = torch.empty(B, n_samples, device=dev, dtype=torch.float32)
    _lib.require_cuda(out)
    if B:
        with torch.cuda.device(dev):
            _lib.check(_lib.load().avhot_synth_audio(_lib.ptr(out), B, n_samples, clip0, seed, _lib.stream_ptr(dev)),
                       "synth_audio")
    return out


def synth_video(B: int, clip0: int, seed: int, device, n_frames: int = N_FRAMES_VIDEO, frame_dim: int = FRAME_DIM,
                dtype=torch.float32) -> torch.Tensor:
    """fp32 k/255 (src/dataprocessing.py:261-265) or the raw uint8 pixels k."""
    dev = torch.device(device)
    out = torch.empty(B, n_frames, frame_dim, device=dev, dtype=dtype)
    _lib.require_cuda(out)
    if dtype not in (torch.float32, torch.uint8):
        raise ValueError("dtype must be float32 or uint8")
    lib = _lib.load()
    with torch.cuda.device(dev):
        for b0 in range(0, B, 32768):
            nb = min(32768, B - b0)
            o = out[b0:b0 + nb]
            _lib.check(lib.avhot_synth_video(_lib.ptr(o) if dtype == torch.float32 else None,
                                             _lib.ptr(o) if dtype == torch.uint8 else None, nb, n_frames * frame_dim,
                                             clip0 + b0, seed, _lib.stream_ptr(dev)), "synth_video")
    return out


def modality_mask(B: int, clip0: int, p: float, seed: int, device, num_modalities: int = 2) -> torch.Tensor:
    """src/data.py:108-118 modality dropout, (B, M) fp32, 1 = available."""
    dev = torch.device(device)
    out = torch.empty(B, num_modalities, device=dev, dtype=torch.float32)
    _lib.require_cuda(out)
    if B:
        with torch.cuda.device(dev):
            _lib.check(_lib.load().avhot_modality_mask(_lib.ptr(out), B, num_modalities, clip0, float(p), seed,
                                                       _lib.stream_ptr(dev)), "modality_mask")
    return out


def dropout_keep(B: int, n_elems: int, clip0: int, t: int, site: int, p: float, seed: int, device) -> torch.Tensor:
    """The keep mask (B, n_elems) uint8 the fused kernels apply at a dropout site (diagnostic / tests)."""
    dev = torch.device(device)
    out = torch.empty(B, n_elems, device=dev, dtype=torch.uint8)
    _lib.require_cuda(out)
    if B:
        with torch.cuda.device(dev):
            _lib.check(_lib.load().avhot_dropout_keep(_lib.ptr(out), B, n_elems, clip0, t, site, float(p), seed,
                                                      _lib.stream_ptr(dev)), "dropout_keep")
    return out
