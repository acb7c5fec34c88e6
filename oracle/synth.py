"""Seeded synthetic RAVDESS-shaped inputs and weights.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Every value is a pure function of (seed, global clip index, element index)
through Philox-4x32-10 (oracle/philox.py), so shards agree at any world size
and csrc/synth.cu reproduces the inputs BIT-EXACTLY on the device (only exact
integer->float conversions, power-of-two scalings, and single correctly-rounded
fp32 multiplies / adds / divides are used - no transcendental functions).

Shapes follow BASELINE.json / SURVEY.md 8(d):
  audio  (B, 48000) fp32  post-conditions of load_raw_audio
         (src/dataprocessing.py:198-211): fixed length, right zero padding,
         divided by max|y| when that is > 0.
  video  (B, 16, 12544) fp32 in {k/255}  (src/dataprocessing.py:261-265).

Streams (Philox counter word c3):
  0x100 audio noise      ctr=(i, g, 0, 0x100): s = sum_j (w_j >> 10)  (4 x 22 bit)
                         n_i = (fp32(s) - 8388606) * 2**-22          (Irwin-Hall(4), exact)
  0x101 audio envelope   block b = i // 800: u = (words(b>>2, g, 0, 0x101)[b&3] >> 8) * 2**-24
                         gain_b = fp32(0.05) + fp32(0.95) * u        (mul, then add, each rounded)
  0x102 audio zero tail  L = words(0, g, 0, 0x102)[0] % 16001 trailing samples are 0
  0x110 video pixels     k = words(i>>2, g, 0, 0x110)[i&3] >> 24;  x = fp32(k) / fp32(255)
  0x120 weights          tensor id = crc32(name): u = (words(e>>2, id, 0, 0x120)[e&3] >> 8) * 2**-24
"""
from __future__ import annotations

import zlib
from collections import OrderedDict

import numpy as np

from . import philox as px

N_SAMPLES = 48000
N_FRAMES_VIDEO = 16
FRAME_DIM = 112 * 112
ENV_BLOCK = 800

STREAM_AUDIO_NOISE = 0x100
STREAM_AUDIO_ENV = 0x101
STREAM_AUDIO_TAIL = 0x102
STREAM_VIDEO = 0x110
STREAM_WEIGHTS = 0x120


def synth_audio(clip_ids, seed: int, n_samples: int = N_SAMPLES) -> np.ndarray:
    clip_ids = np.asarray(clip_ids, dtype=np.uint64)
    B = clip_ids.shape[0]
    k0, k1 = px.seed_key(seed)
    out = np.empty((B, n_samples), dtype=np.float32)
    i = np.arange(n_samples, dtype=np.uint64)
    nblk = (n_samples + ENV_BLOCK - 1) // ENV_BLOCK
    for b in range(B):
        g = clip_ids[b]
        w = px.philox4x32_10(i, g, 0, STREAM_AUDIO_NOISE, k0, k1)
        s = sum((x >> np.uint32(10)).astype(np.int64) for x in w)
        noise = (s.astype(np.float32) - np.float32(8388606.0)) * np.float32(2.0 ** -22)
        ew = px.element_words(nblk, g, 0, STREAM_AUDIO_ENV, seed)
        u = (ew >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
        gain = np.float32(0.05) + np.float32(0.95) * u
        y = noise * np.repeat(gain, ENV_BLOCK)[:n_samples]
        tail = int(px.philox4x32_10(0, g, 0, STREAM_AUDIO_TAIL, k0, k1)[0]) % 16001
        tail = min(tail, n_samples)
        if tail:
            y[n_samples - tail:] = 0.0
        peak = np.max(np.abs(y))
        if peak > 0:
            y = y / peak
        out[b] = y
    return out


def synth_video_u8(clip_ids, seed: int, n_frames: int = N_FRAMES_VIDEO,
                   frame_dim: int = FRAME_DIM) -> np.ndarray:
    clip_ids = np.asarray(clip_ids, dtype=np.uint64)
    n = n_frames * frame_dim
    w = px.element_words(n, clip_ids, 0, STREAM_VIDEO, seed)
    return (w >> np.uint32(24)).astype(np.uint8).reshape(clip_ids.shape[0], n_frames, frame_dim)


def synth_video(clip_ids, seed: int, n_frames: int = N_FRAMES_VIDEO,
                frame_dim: int = FRAME_DIM) -> np.ndarray:
    k = synth_video_u8(clip_ids, seed, n_frames, frame_dim)
    return k.astype(np.float32) / np.float32(255.0)


# --------------------------------------------------------------------------
# Weights.  Key names/shapes = SURVEY.md Appendix A (dumped from the reference
# classes); tests/golden/make_golden.py loads them into the reference modules
# with strict=True, which proves the layout.
# --------------------------------------------------------------------------

def _uniform(name: str, shape, seed: int) -> np.ndarray:
    n = int(np.prod(shape))
    tid = zlib.crc32(name.encode()) & 0xFFFFFFFF
    w = px.element_words(n, np.uint64(tid), 0, STREAM_WEIGHTS, seed)
    u = (w >> np.uint32(8)).astype(np.float64) * 2.0 ** -24
    return (2.0 * u - 1.0).reshape(shape)


def model_shapes(audio_dim=104, frame_dim=FRAME_DIM, enc_hidden=256, out_dim=128,
                 fusion_hidden=256, num_classes=8, lstm_layers=2,
                 modalities=("audio", "video")) -> "OrderedDict[str, tuple]":
    """(name -> (shape, kind, fan_in)) for encoders + HybridFusion + baseline head."""
    s = OrderedDict()
    H = enc_hidden
    for l in range(lstm_layers):
        ind = audio_dim if l == 0 else H
        s[f"encoders.audio.rnn.weight_ih_l{l}"] = ((4 * H, ind), "u", H)
        s[f"encoders.audio.rnn.weight_hh_l{l}"] = ((4 * H, H), "u", H)
        s[f"encoders.audio.rnn.bias_ih_l{l}"] = ((4 * H,), "u", H)
        s[f"encoders.audio.rnn.bias_hh_l{l}"] = ((4 * H,), "u", H)
    s["encoders.audio.projection.weight"] = ((out_dim, H), "u", H)
    s["encoders.audio.projection.bias"] = ((out_dim,), "u", H)
    # SequenceEncoder(encoder_type='gru') at the same dims (src/encoders.py:65-77), an alternative audio encoder kept
    # under its own prefix and loaded as a standalone module by its consumers
    for l in range(lstm_layers):
        ind = audio_dim if l == 0 else H
        s[f"gru_enc.rnn.weight_ih_l{l}"] = ((3 * H, ind), "u", H)
        s[f"gru_enc.rnn.weight_hh_l{l}"] = ((3 * H, H), "u", H)
        s[f"gru_enc.rnn.bias_ih_l{l}"] = ((3 * H,), "u", H)
        s[f"gru_enc.rnn.bias_hh_l{l}"] = ((3 * H,), "u", H)
    s["gru_enc.projection.weight"] = ((out_dim, H), "u", H)
    s["gru_enc.projection.bias"] = ((out_dim,), "u", H)
    # SequenceEncoder(encoder_type='cnn') (src/encoders.py:86-93); BatchNorm running statistics perturbed like the
    # LayerNorm affines so the folded arithmetic is exercised (running_var = 1 + 0.1 u > 0)
    s["cnn_enc.conv1.weight"] = ((H, audio_dim, 5), "u", 5 * audio_dim)
    s["cnn_enc.conv1.bias"] = ((H,), "u", 5 * audio_dim)
    s["cnn_enc.conv2.weight"] = ((H, H, 3), "u", 3 * H)
    s["cnn_enc.conv2.bias"] = ((H,), "u", 3 * H)
    for bn in ("bn1", "bn2"):
        s[f"cnn_enc.{bn}.weight"] = ((H,), "lnw", 0)
        s[f"cnn_enc.{bn}.bias"] = ((H,), "lnb", 0)
        s[f"cnn_enc.{bn}.running_mean"] = ((H,), "lnb", 0)
        s[f"cnn_enc.{bn}.running_var"] = ((H,), "lnw", 0)
    s["cnn_enc.projection.weight"] = ((out_dim, H), "u", H)
    s["cnn_enc.projection.bias"] = ((out_dim,), "u", H)
    s["encoders.video.frame_mlp.0.weight"] = ((H, frame_dim), "u", frame_dim)
    s["encoders.video.frame_mlp.0.bias"] = ((H,), "u", frame_dim)
    s["encoders.video.attention.weight"] = ((1, H), "u", H)
    s["encoders.video.attention.bias"] = ((1,), "u", H)
    s["encoders.video.projection.0.weight"] = ((H,), "lnw", 0)
    s["encoders.video.projection.0.bias"] = ((H,), "lnb", 0)
    s["encoders.video.projection.1.weight"] = ((out_dim, H), "u", H)
    s["encoders.video.projection.1.bias"] = ((out_dim,), "u", H)
    D = fusion_hidden
    G = max(32, D // 2)
    for m in modalities:
        s[f"fusion_model.proj_layers.{m}.weight"] = ((D, out_dim), "u", out_dim)
        s[f"fusion_model.proj_layers.{m}.bias"] = ((D,), "u", out_dim)
    for m in modalities:
        p = f"fusion_model.attn_layers.{m}."
        for ln in ("q_in_ln", "k_in_ln", "v_in_ln"):
            s[p + ln + ".weight"] = ((D,), "lnw", 0)
            s[p + ln + ".bias"] = ((D,), "lnb", 0)
        for pr in ("q_proj", "k_proj", "v_proj", "out_proj"):
            s[p + pr + ".weight"] = ((D, D), "u", D)
            s[p + pr + ".bias"] = ((D,), "u", D)
    for ln in ("pre_ln", "post_ln"):
        s[f"fusion_model.{ln}.weight"] = ((D,), "lnw", 0)
        s[f"fusion_model.{ln}.bias"] = ((D,), "lnb", 0)
    s["fusion_model.fusion_gate.0.weight"] = ((G, D), "u", D)
    s["fusion_model.fusion_gate.0.bias"] = ((G,), "u", D)
    s["fusion_model.fusion_gate.3.weight"] = ((1, G), "u", G)
    s["fusion_model.fusion_gate.3.bias"] = ((1,), "u", G)
    s["fusion_model.classifier.weight"] = ((num_classes, D), "u", D)
    s["fusion_model.classifier.bias"] = ((num_classes,), "u", D)
    cat = out_dim * len(modalities)
    # EarlyFusion (src/fusion.py:60-72) and LateFusion (:180-189) at the same dims; they are alternatives to
    # fusion_model, kept under their own prefixes here and loaded as `fusion_model.*` by the consumers
    s["early_model.mlp.0.weight"] = ((D, cat), "u", cat)
    s["early_model.mlp.0.bias"] = ((D,), "u", cat)
    s["early_model.mlp.1.weight"] = ((D,), "lnw", 0)
    s["early_model.mlp.1.bias"] = ((D,), "lnb", 0)
    s["early_model.mlp.4.weight"] = ((D, D), "u", D)
    s["early_model.mlp.4.bias"] = ((D,), "u", D)
    s["early_model.mlp.5.weight"] = ((D,), "lnw", 0)
    s["early_model.mlp.5.bias"] = ((D,), "lnb", 0)
    s["early_model.mlp.8.weight"] = ((num_classes, D), "u", D)
    s["early_model.mlp.8.bias"] = ((num_classes,), "u", D)
    for m in modalities:
        s[f"late_model.classifiers.{m}.0.weight"] = ((D, out_dim), "u", out_dim)
        s[f"late_model.classifiers.{m}.0.bias"] = ((D,), "u", out_dim)
        s[f"late_model.classifiers.{m}.3.weight"] = ((num_classes, D), "u", D)
        s[f"late_model.classifiers.{m}.3.bias"] = ((num_classes,), "u", D)
    s["late_model.fusion_logits"] = ((len(modalities),), "lnb", 0)
    # LateFusionWithUncertainty (src/fusion.py:611-620) with hidden_dim = D
    for m in modalities:
        for head, outs in (("classifiers", num_classes), ("uncert_heads", 1)):
            s[f"unc_model.{head}.{m}.1.weight"] = ((D, out_dim), "u", out_dim)
            s[f"unc_model.{head}.{m}.1.bias"] = ((D,), "u", out_dim)
            s[f"unc_model.{head}.{m}.4.weight"] = ((outs, D), "u", D)
            s[f"unc_model.{head}.{m}.4.bias"] = ((outs,), "u", D)
    s["fusion_head.0.weight"] = ((D, cat), "u", cat)
    s["fusion_head.0.bias"] = ((D,), "u", cat)
    s["fusion_head.2.weight"] = ((num_classes, D), "u", D)
    s["fusion_head.2.bias"] = ((num_classes,), "u", D)
    return s


def synth_weights(seed: int, gain: float = 1.0, **shape_kwargs) -> "OrderedDict[str, np.ndarray]":
    """torch-default-style init, U(-1/sqrt(fan_in), +1/sqrt(fan_in)) for Linear/LSTM
    tensors (what nn.Linear / nn.LSTM reset_parameters draw from); LayerNorm
    affine is perturbed (1 + 0.1u, 0.1u) instead of the default (1, 0) so that
    tests exercise the affine terms.  `gain` scales the non-LN tensors (gain > 1
    gives better-separated logits than an untrained model)."""
    out = OrderedDict()
    for name, (shape, kind, fan_in) in model_shapes(**shape_kwargs).items():
        u = _uniform(name, shape, seed)
        if kind == "u":
            v = u * (gain / np.sqrt(fan_in))
        elif kind == "lnw":
            v = 1.0 + 0.1 * u
        else:
            v = 0.1 * u
        out[name] = v.astype(np.float32)
    return out


def state_dict_for(head: str, sd):
    """The tensors a model with the given head holds, under the reference's key names:
    encoders.* plus fusion_model.* (hybrid / early / late: the three alternatives of build_fusion_model share the
    attribute name, src/train.py:69) or fusion_head.* (concat, the live path)."""
    src = {"hybrid": "fusion_model.", "early": "early_model.", "late": "late_model.", "concat": "fusion_head."}[head]
    dst = "fusion_head." if head == "concat" else "fusion_model."
    out = OrderedDict()
    for k, v in sd.items():
        if k.startswith("encoders."):
            out[k] = v
        elif k.startswith(src):
            out[dst + k[len(src):]] = v
    return out
