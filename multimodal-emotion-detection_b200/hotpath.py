"""The per-clip audio-visual inference hot path as one module.

AVHotPath mirrors MultimodalFusionModule (src/train.py:34-164) without Lightning: ``encoders`` (ModuleDict, in
``modalities`` order, :55-66), then either ``fusion_model`` = HybridFusion / EarlyFusion / LateFusion (the commented
composition :69-76, :138-147, `model.fusion_type`) or ``fusion_head`` = Linear-ReLU-Linear on the concatenation (:81-85, :153-164, the live path, which
ignores the mask) - same attribute names, so state-dict keys match SURVEY.md Appendix A.

forward(features, mask)   one pass, like the reference's forward.
mc_forward(features, mask, num_samples, ...)   MCDropoutUncertainty's result with the pass-invariant stages
                          (front-end, LSTM, video Linear + ReLU) computed once and the T stochastic passes inside
                          the fused tail kernels.

features["audio"] may be the raw waveform (B, 48000) / (B, 48000, 1) as load_raw_audio produces it
(src/dataprocessing.py:210-211) - the log-mel/MFCC front-end is applied -, the same as 16-bit PCM (int16, as the .wav
holds it: load_raw_audio's peak normalisation then happens on the device), or precomputed (B, T, 104) features.
features["video"] is (B, 16, 12544) fp32 in [0, 1] or uint8 pixels.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib, _rng, dataprocessing
from .encoders import build_encoder
from .fusion import ConcatHead, build_fusion_model
from .uncertainty import mc_reduce

DEFAULT_ENCODERS = {   # configs/base.yaml:30-47 with BASELINE.json's input dims
    "audio": {"type": "sequence", "input_dim": dataprocessing.N_FEATS, "encoder_type": "lstm", "hidden_dim": 256,
              "num_layers": 2, "dropout": 0.1},
    "video": {"type": "frame", "input_dim": 112 * 112, "temporal_pooling": "attention", "hidden_dim": 256,
              "dropout": 0.1},
}


class AVHotPath(nn.Module):
    def __init__(self, head: str = "hybrid", modalities=("audio", "video"), encoders: Optional[dict] = None,
                 output_dim: int = 128, hidden_dim: int = 256, num_heads: int = 4, num_classes: int = 8,
                 dropout: float = 0.3, precision: str = "fp32"):
        super().__init__()
        if head not in ("hybrid", "concat", "early", "late"):
            raise ValueError(f"Unknown head: {head}")
        self.head = head
        self.modalities = list(modalities)
        cfgs = encoders or DEFAULT_ENCODERS
        self.encoders = nn.ModuleDict()
        dims = {}
        for m in self.modalities:                                                        # src/train.py:55-66
            cfg = cfgs[m]
            self.encoders[m] = build_encoder(modality=m, input_dim=cfg.get("input_dim", 64), output_dim=output_dim,
                                             encoder_config=cfg)
            dims[m] = output_dim
        if head != "concat":                                                             # hybrid / early / late
            self.fusion_model = build_fusion_model(head, dims, num_classes, hidden_dim=hidden_dim,
                                                   num_heads=num_heads, dropout=dropout)  # :69-76
        else:
            self.fusion_head = ConcatHead(output_dim * len(self.modalities), hidden_dim, num_classes,
                                          len(self.modalities))                          # :81-85
        self.num_classes = num_classes
        self.fold_video_projection = True     # tc + hybrid only; False keeps z_v materialised (A/B tests)
        self._resolver = None                 # exact.NearTieResolver of mc_forward's immediate mode
        self.near_tie_margin_rel = None       # None = exact.DEFAULT_MARGIN_REL (units of the model's logit scale)
        self.set_precision(precision)

    def set_precision(self, precision: str) -> "AVHotPath":
        """'fp32': every contraction in fp32 FMA (the 1e-4 parity path).  'tc': the large contractions (video
        Linear, LSTM input projections and recurrence) on the tcgen05 tensor cores (TF32 / fp16 operands, fp32
        accumulate; tolerance stated in DESIGN.md), and so are the per-pass GEMMs of HybridFusion (fp16 operands)."""
        if precision not in ("fp32", "tc"):
            raise ValueError(f"Unknown precision: {precision}")
        self.precision = precision
        code = _lib.PREC_FP32 if precision == "fp32" else _lib.PREC_TC
        for enc in self.encoders.values():
            enc.precision = code
        if self.head == "hybrid":
            self.fusion_model.precision = code
        return self

    # ---- stages ----------------------------------------------------------------------------------------
    def audio_features(self, audio: torch.Tensor) -> torch.Tensor:
        enc = self.encoders["audio"]
        if audio.dim() == 3 and audio.shape[-1] == enc.input_dim and enc.input_dim != 1:
            return audio                                   # already (B, T, 104)
        return dataprocessing.logmel_mfcc(audio)           # raw waveform (B, N) or (B, N, 1)

    @torch.no_grad()
    def _run(self, features: Dict[str, torch.Tensor], mask, passes: int, stochastic: bool, seed: int, clip0: int,
             t0: int, clip_ids=None):
        """clip_ids: optional (B,) int32 CUDA tensor, the global clip index of every row (instead of clip0 + b); used by
        exact.NearTieResolver to re-evaluate scattered clips on the fp32 kernels (hybrid / concat heads)."""
        if set(self.modalities) != {"audio", "video"}:
            raise ValueError("AVHotPath is the audio + video path")
        audio, video = features["audio"], features["video"]
        dev = _lib.require_cuda(audio, video)
        _lib.check_device(dev)
        if audio.dtype == torch.int16:        # 16-bit PCM as decoded from the .wav: load_raw_audio's conditioning here
            if audio.dim() == 3 and audio.shape[-1] == 1:
                audio = audio[..., 0]
            audio = dataprocessing.condition_audio(audio)
        B = audio.shape[0]
        aenc = self.encoders["audio"]
        raw = not (audio.dim() == 3 and audio.shape[-1] == aenc.input_dim and aenc.input_dim != 1)
        if raw and self.precision == "tc" and B <= 65535 and aenc.encoder_type == "lstm":
            z_a = aenc.forward_from_waveform(audio)           # front-end fused with the LSTM operand preparation
        else:
            z_a = aenc(self.audio_features(audio))                                        # (B, out), pass-invariant
        venc = self.encoders["video"]
        hid = venc.hidden(video)                                                          # (B, F, H), pass-invariant
        p_enc = venc._drop_p() if stochastic else 0.0
        # tensor-core hybrid path: the frame encoder stops after its LayerNorm and its 256 -> 128 projection is folded
        # into HybridFusion's proj_layers.video (two linear maps back to back), which is a tcgen05 GEMM anyway
        fold = (self.head == "hybrid" and self.precision == "tc" and self.modalities[0] == "audio"
                and venc.hidden_dim % 64 == 0 and venc.hidden_dim <= 256 and self.fold_video_projection)
        if clip_ids is not None and (fold or self.head not in ("hybrid", "concat")):
            raise ValueError("clip_ids: only the fp32 hybrid / concat paths take per-row clip indices")
        z_v = venc.tail(hid, None, passes, p_enc, seed, clip0, t0, project=not fold,
                        clip_ids=clip_ids)                                                # (B, passes, out | hidden)
        zs = [z_a, z_v.reshape(B * passes, -1)] if self.modalities[0] == "audio" else [z_v.reshape(B * passes, -1), z_a]
        rpc = [1, passes] if self.modalities[0] == "audio" else [passes, 1]
        if self.head == "hybrid":
            fm = self.fusion_model
            mask_f = None
            if mask is not None:
                mask_f = torch.as_tensor(mask).to(device=dev, dtype=torch.float32).contiguous()
                if mask_f.shape != (B, 2):
                    raise ValueError("mask must be (B, 2)")
            p_f = fm._drop_p() if stochastic else 0.0
            logits, _, _, _ = fm.run(zs, rpc, mask_f, B, passes, p_f, seed, clip0, t0,
                                     folded_from=(1, venc.projection[1]) if fold else None, clip_ids=clip_ids)
        elif self.head in ("early", "late"):
            fm = self.fusion_model
            mask_f = None
            if mask is not None:
                mask_f = torch.as_tensor(mask).to(device=dev, dtype=torch.float32).contiguous()
                if mask_f.shape != (B, 2):
                    raise ValueError("mask must be (B, 2)")
            logits, _ = fm.run(zs, rpc, mask_f, B, passes, fm._drop_p() if stochastic else 0.0, seed, clip0, t0)
        else:
            logits = self.fusion_head.run(zs, rpc, B, passes)                             # mask ignored (F2)
        return logits                                                                     # (B, passes, C)

    def forward(self, features: Dict[str, torch.Tensor], mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        seed, t, clip0 = _rng.current()
        return self._run(features, mask, 1, True, seed, clip0, t)[:, 0, :]

    def mc_forward(self, features: Dict[str, torch.Tensor], mask: Optional[torch.Tensor] = None,
                   num_samples: int = 32, seed: int = _rng.DEFAULT_SEED, clip0: int = 0, exact_labels=None):
        """-> (mean_logits (B, C), uncertainty (B,), labels (B,) int32, packed (B, C + 1)).  Dropout modules must
        be in the mode the caller wants (MCDropoutUncertainty puts exactly the nn.Dropout modules in train mode).

        exact_labels (precision 'tc', hybrid / concat heads): the clips whose top-2 margin of the mean logits is inside
        the tensor-core path's stated tolerance are re-evaluated on the fp32 kernels, so every label equals the fp32
        path's (exact.py).  None / True: before returning (one host round trip per call); a NearTieResolver: deferred -
        the batch is registered and the caller flushes the resolver every few batches; False: off."""
        logits = self._run(features, mask, num_samples, True, seed, clip0, 0)
        results = mc_reduce(logits)
        if exact_labels is False or self.precision != "tc" or self.head not in ("hybrid", "concat"):
            return results
        from .exact import NearTieResolver
        if isinstance(exact_labels, NearTieResolver):
            exact_labels.register(features, mask, results, num_samples, seed, clip0)
            return results
        B = results[0].shape[0]
        want = B if B <= 256 else max(256, B // 8)             # near ties are a fraction of a percent of a batch
        from .exact import DEFAULT_MARGIN_REL
        rel = DEFAULT_MARGIN_REL if self.near_tie_margin_rel is None else float(self.near_tie_margin_rel)
        if self._resolver is None or self._resolver.capacity < want or self._resolver.margin_rel != rel:
            self._resolver = NearTieResolver(self, capacity=want, margin_rel=rel)
        self._resolver.register(features, mask, results, num_samples, seed, clip0)
        if self._resolver.flush(raise_on_overflow=False) < 0:
            # a degenerate model (e.g. all-equal logits) puts most clips inside the tolerance: the fp32 path for all
            self.set_precision("fp32")
            try:
                results = mc_reduce(self._run(features, mask, num_samples, True, seed, clip0, 0))
            finally:
                self.set_precision("tc")
        return results
