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

from typing import Tuple

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
        self.seed = _rng.DEFAULT_SEED
        self.clip0 = 0

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
        return mean, unc
