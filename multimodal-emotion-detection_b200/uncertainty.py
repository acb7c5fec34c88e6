"""Drop-in for src/uncertainty.py: MCDropoutUncertainty (src/uncertainty.py:17-86).

Semantics kept: model.eval(), then only nn.Dropout* modules in train mode (:36-46, :58-59); num_samples
stochastic forwards (:62-65); mean of LOGITS over passes (:77); population variance of softmax PROBABILITIES
over passes, averaged over classes (:80-84); the model's train/eval mode is restored (:67-71).

Two execution paths, same results:
* any module built from this package: the T passes run one after the other, pass t drawing the Philox masks of
  (seed, t) (see _rng.py), and the statistics are reduced by avhot_mc_reduce;
* a model that exposes ``mc_forward`` (AVHotPath): the pass-invariant work (front-end, audio encoder, video
  Linear + ReLU) is done once and the T passes run inside the fused tail kernels.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn as nn

from . import _lib, _rng


def mc_reduce(logits: torch.Tensor):
    """logits (B, T, C) CUDA fp32 -> (mean_logits (B, C), uncertainty (B,), labels (B,) int32,
    packed (B, C + 1) = [mean_logits | uncertainty])."""
    dev = _lib.require_cuda(logits)
    _lib.check_device(dev)
    logits = _lib.f32c(logits)
    B, T, Cn = logits.shape
    mean = torch.empty(B, Cn, device=dev, dtype=torch.float32)
    unc = torch.empty(B, device=dev, dtype=torch.float32)
    labels = torch.empty(B, device=dev, dtype=torch.int32)
    packed = torch.empty(B, Cn + 1, device=dev, dtype=torch.float32)
    if B > 0:
        with torch.cuda.device(dev):
            _lib.check(_lib.load().avhot_mc_reduce(_lib.ptr(logits), _lib.ptr(mean), _lib.ptr(unc), _lib.ptr(labels),
                                                   _lib.ptr(packed), B, T, Cn, _lib.stream_ptr(dev)), "mc_reduce")
    return mean, unc, labels, packed


class MCDropoutUncertainty(nn.Module):
    def __init__(self, model: nn.Module, num_samples: int = 10):
        super().__init__()
        assert num_samples >= 1, "num_samples must be >= 1"                               # :31
        self.model = model
        self.num_samples = num_samples
        # Dropout masks are a pure function of (seed, pass, global clip index): row b of a call is clip `clip0 + b`.
        # clip0 advances by the batch size after every call, so consecutive batches of a loader draw fresh masks like
        # the reference's nn.Dropout does (src/uncertainty.py:61-65); set it (or advance = False) to pin the masks.
        self.seed = _rng.DEFAULT_SEED
        self.clip0 = 0
        self.advance = True

    @staticmethod
    def _set_dropout_train(module: nn.Module) -> None:                                    # :36-46
        dropout_types = (nn.Dropout, nn.Dropout1d, nn.Dropout2d, nn.Dropout3d, nn.AlphaDropout)
        for m in module.modules():
            if isinstance(m, dropout_types):
                m.train()

    def forward(self, *args, **kwargs) -> Tuple[torch.Tensor, torch.Tensor]:
        was_training = self.model.training
        self.model.eval()
        self._set_dropout_train(self.model)
        try:
            if hasattr(self.model, "mc_forward"):
                mean, unc, _, _ = self.model.mc_forward(*args, num_samples=self.num_samples, seed=self.seed,
                                                        clip0=self.clip0, **kwargs)
            else:
                samples = []
                for t in range(self.num_samples):
                    with _rng.philox_context(self.seed, t, self.clip0):
                        out = self.model(*args, **kwargs)
                    samples.append(out[0] if isinstance(out, tuple) else out)
                mean, unc, _, _ = mc_reduce(torch.stack(samples, dim=1))
        finally:
            if was_training:
                self.model.train()
            else:
                self.model.eval()
        if self.advance:
            self.clip0 += int(mean.shape[0])
        return mean, unc


class UncertaintyWeightedFusion(nn.Module):
    """src/uncertainty.py:279-318: fuse per-modality logits weighted by inverse uncertainty,
    w_m = (mask_m / (u_m + eps)) / (sum_m + eps).  One small kernel (avhot_uncertainty_fuse)."""

    def __init__(self, epsilon: float = 1e-6):
        super().__init__()
        self.epsilon = epsilon

    def forward(self, modality_predictions: Dict[str, torch.Tensor], modality_uncertainties: Dict[str, torch.Tensor],
                modality_mask: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        assert isinstance(modality_predictions, dict) and isinstance(modality_uncertainties, dict)      # :301
        modalities = list(modality_predictions.keys())
        logits = _lib.f32c(torch.stack([modality_predictions[m] for m in modalities], dim=1))           # (B, M, C)
        dev = _lib.require_cuda(logits)
        _lib.check_device(dev)
        uncert = _lib.f32c(torch.stack([modality_uncertainties[m] for m in modalities], dim=1).to(dev))
        mask = _lib.f32c(modality_mask.to(device=dev))
        B, M, Cn = logits.shape
        if uncert.shape != (B, M) or mask.shape != (B, M):
            raise ValueError(f"uncertainties and mask must be (B, {M})")
        fused = torch.empty(B, Cn, device=dev, dtype=torch.float32)
        weights = torch.empty(B, M, device=dev, dtype=torch.float32)
        if B > 0:
            with torch.cuda.device(dev):
                _lib.check(_lib.load().avhot_uncertainty_fuse(_lib.ptr(logits), _lib.ptr(uncert), _lib.ptr(mask),
                                                              _lib.ptr(fused), _lib.ptr(weights), B, M, Cn,
                                                              float(self.epsilon), _lib.stream_ptr(dev)),
                           "UncertaintyWeightedFusion")
        return fused, weights


# ---- calibration metrics (src/uncertainty.py:88-183, :471-544) on the GPU ---------------------
def calibration_accumulate(acc: torch.Tensor, labels: torch.Tensor, num_bins: int, logits: torch.Tensor = None,
                           confidences: torch.Tensor = None, predictions: torch.Tensor = None) -> torch.Tensor:
    """Adds one batch to the (3 * num_bins + 3,) float64 CUDA accumulator of avhot_calibration_accumulate:
    [bin counts | sum of confidences | correct counts | NLL sum, correct total, samples]."""
    dev = _lib.require_cuda(acc, labels, logits, confidences, predictions)
    _lib.check_device(dev)
    labels = labels.long().contiguous()
    n = labels.shape[0]
    edges = torch.linspace(0, 1, steps=num_bins + 1, device=dev)                          # :118
    cn = 0
    if logits is not None:
        logits = _lib.f32c(logits)
        assert logits.ndim == 2 and logits.shape[0] == n, "logits (N,C), labels (N,)"
        cn = logits.shape[1]
    else:
        confidences = _lib.f32c(confidences.detach())
        predictions = predictions.detach().long().contiguous()
    if n > 0:
        with torch.cuda.device(dev):
            _lib.check(_lib.load().avhot_calibration_accumulate(
                _lib.ptr(logits), _lib.ptr(confidences), _lib.ptr(predictions), _lib.ptr(labels), n, cn,
                _lib.ptr(edges), num_bins, _lib.ptr(acc), _lib.stream_ptr(dev)), "calibration")
    return acc


def calibration_finalize(acc: torch.Tensor, num_bins: int) -> Dict[str, float]:
    """ECE / MCE / NLL / accuracy from the accumulator, exactly as the reference combines its per-bin statistics
    (empty bins skipped, :132-152; ECE :163-171; MCE :183; NLL = sum / max(1, N), :533)."""
    a = acc.detach().cpu().double()
    cnt, conf, cor = a[:num_bins], a[num_bins:2 * num_bins], a[2 * num_bins:3 * num_bins]
    nll_sum, cor_tot, n = float(a[3 * num_bins]), float(a[3 * num_bins + 1]), float(a[3 * num_bins + 2])
    tot = float(cnt.sum())
    ece, mce = 0.0, 0.0
    for b in range(num_bins):
        if cnt[b] > 0:
            gap = abs(float(cor[b] / cnt[b]) - float(conf[b] / cnt[b]))
            ece += gap * float(cnt[b]) / tot
            mce = max(mce, gap)
    return {"ece": ece, "mce": mce, "nll": nll_sum / max(1.0, n), "accuracy": cor_tot / n if n else 0.0}


class CalibrationMetrics:
    """src/uncertainty.py:88-202 with the binning done by one CUDA histogram kernel (inputs must be CUDA tensors)."""

    @staticmethod
    def _acc(confidences, predictions, labels, num_bins):
        assert confidences.ndim == predictions.ndim == labels.ndim == 1, "Use 1D tensors"             # :108
        assert confidences.shape[0] == predictions.shape[0] == labels.shape[0], "Mismatched shapes"   # :109
        acc = torch.zeros(3 * num_bins + 3, device=confidences.device, dtype=torch.float64)
        return calibration_accumulate(acc, labels, num_bins, confidences=confidences, predictions=predictions)

    @staticmethod
    def expected_calibration_error(confidences, predictions, labels, num_bins: int = 15) -> float:
        return calibration_finalize(CalibrationMetrics._acc(confidences, predictions, labels, num_bins), num_bins)["ece"]

    @staticmethod
    def maximum_calibration_error(confidences, predictions, labels, num_bins: int = 15) -> float:
        return calibration_finalize(CalibrationMetrics._acc(confidences, predictions, labels, num_bins), num_bins)["mce"]

    @staticmethod
    def negative_log_likelihood(logits: torch.Tensor, labels: torch.Tensor) -> float:
        assert logits.ndim == 2 and labels.ndim == 1, "logits (N,C), labels (N,)"                      # :197
        acc = torch.zeros(3 + 3, device=logits.device, dtype=torch.float64)
        return calibration_finalize(calibration_accumulate(acc, labels, 1, logits=logits), 1)["nll"]


def compute_calibration_metrics(model: nn.Module, dataloader, device: str = "cuda", num_bins: int = 15) -> Dict[str, float]:
    """src/uncertainty.py:471-544: ECE, MCE, NLL, accuracy over a dataloader of (inputs, labels) or
    (features_dict, labels, mask) batches; models returning (logits, aux) are unwrapped.  The per-batch statistics
    stay on the GPU (one histogram kernel per batch into a float64 accumulator); one D2H copy at the end."""
    model.eval().to(device)
    acc = None
    for batch in dataloader:
        if isinstance(batch, (tuple, list)):
            if len(batch) == 2:
                inputs, labels = batch
                mask = None
            elif len(batch) == 3:
                inputs, labels, mask = batch
            else:
                raise ValueError(f"Unexpected batch shape: len={len(batch)}")
        else:
            raise ValueError("Batch must be a tuple/list")
        labels = labels.to(device)
        if isinstance(inputs, dict):
            inputs = {k: v.to(device) for k, v in inputs.items()}
            mask = mask.to(device) if mask is not None else None
            logits = model(inputs, mask)
        else:
            logits = model(inputs.to(device))
        if isinstance(logits, (tuple, list)):
            logits = logits[0]
        if acc is None:
            acc = torch.zeros(3 * num_bins + 3, device=logits.device, dtype=torch.float64)
        calibration_accumulate(acc, labels, num_bins, logits=logits)
    if acc is None:
        return {"ece": 0.0, "mce": 0.0, "nll": 0.0, "accuracy": 0.0}
    return calibration_finalize(acc, num_bins)
