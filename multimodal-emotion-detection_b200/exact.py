"""Label exactness of the tensor-core path.

BASELINE.json's north star asks for "predicted labels identical to the reference" - the reference predicts
argmax(mean_logits) of its fp32 forward (src/uncertainty.py:77).  A label of the tensor-core path can differ from the
fp32 path's only for a clip whose top-2 margin of the mean logits is smaller than the error of that difference.
The scale of that error is a property of the model, not of the batch: the classifier sees a convex combination of
LayerNorm outputs, so |logit_c - bias_c| <= S = max_c ||w_c * g||_2 sqrt(D) (w_c: classifier row, g: post_ln weight;
S = 9.5 for the default init, where the largest |logit| of a batch is ~1.27).  Measured on 56 832 clips of the benchmark
workload (tools/measure_near_ties.py, profiles/round2/near_ties_margin_measurement.json), in units of S: the error of the top-2 difference has
RMS 8.8e-6, 99.99 % quantile 5.9e-5, maximum 7.1e-5 (the per-pass rounding of the fusion GEMMs averages out over the T
passes; what remains is the pass-invariant TF32 video Linear and the fp16 LSTM state) and a tail that falls a decade
per 1.6e-5; the widest margin of a clip whose label actually flipped was 2.2e-5.  NearTieResolver takes every clip
whose margin is below ``margin_rel * S`` (default 8e-5: an undetected flip then needs an error beyond the measured
tail, ~3e-9 per clip; a pure function of the clip, so results do not depend on batching or sharding), keeps a copy of its
inputs (avhot_flag_near_ties, avhot_gather_flagged: on the device, no host round trip), re-evaluates it through the fp32
kernels with its own global clip index - so the Philox masks are the ones it drew the first time - and overwrites its
records (avhot_scatter_resolved).  With the random-init weights of the benchmark the logits of all clips sit close
together (median top-2 margin 4 % of max|logit|) and 0.83 % of the clips take this path; a trained model's margins are
wider.  The fp32 kernels cost ~17 x the tensor-core path per clip, which is why the threshold is the measured envelope
and not the (looser) tolerance the tests assert on every logit.  bench.py checks the premise on every timed step
(labels and mean logits of a rotating sample against the fp32 path, all ranks).

Two ways to use it:
* immediate (AVHotPath.mc_forward's default with precision 'tc'): flag, re-evaluate and patch before returning;
* deferred (bench.py, sweeps): ``register`` every batch, ``flush`` every few batches - the fp32 recurrence is
  latency-bound below a few hundred clips, so one flush over ten batches costs far less than ten flushes.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from . import _lib

DEFAULT_MARGIN_REL = 8e-5                # of the model's logit scale S; see the measurement above


def logit_scale(model) -> float:
    """S = max_c ||w_c * g||_2 * sqrt(fan_in) of the head's last Linear (g: the LayerNorm gain in front of it, if any):
    an a-priori bound of |logit - bias| for HybridFusion (the classifier's input is a convex combination of post_ln
    rows, src/fusion.py:502-522); for the concat head (Linear-ReLU-Linear, src/train.py:81-85) the same expression on
    fusion_head[2] is the scale the margin is quoted in."""
    if getattr(model, "head", None) == "hybrid":
        lin, gain = model.fusion_model.classifier, model.fusion_model.post_ln.weight
    elif getattr(model, "head", None) == "concat":
        lin, gain = model.fusion_head[2], None
    else:
        raise ValueError("NearTieResolver: hybrid / concat heads only")
    w = lin.weight.detach().float().cpu().numpy().astype("float64")     # a scalar of the model, derived on the host
    if gain is not None:
        w = w * gain.detach().float().cpu().numpy().astype("float64")
    return float((w ** 2).sum(axis=1).max() ** 0.5 * w.shape[1] ** 0.5)


class NearTieResolver:
    def __init__(self, model, capacity: int = 4096, margin_rel: float = DEFAULT_MARGIN_REL):
        if capacity < 1:
            raise ValueError("capacity must be >= 1")
        self.model = model
        self.capacity = int(capacity)
        self.margin_rel = float(margin_rel)
        self.margin = self.margin_rel * logit_scale(model)        # absolute; weights are read once, at construction
        self._dev = None
        self._state = None               # uint32[4]: appended, overflow, appended before the latest call, (unused)
        self._entry = None               # int32 (3, capacity): tag, row, global clip index (uint32 bit patterns)
        self._pool: Dict[str, torch.Tensor] = {}
        self._calls: List[tuple] = []    # (tag, mean, unc, labels, packed)
        self._params = None              # (num_samples, seed) of the pending window
        self._next_tag = 0
        self.flagged_total = 0           # clips re-evaluated so far
        self.clips_total = 0             # clips registered so far

    # -- buffers ---------------------------------------------------------------------------------------------
    def _ensure(self, dev, rows: Dict[str, torch.Tensor]):
        if self._dev is not None and self._dev != dev:
            raise RuntimeError("NearTieResolver is bound to one device")
        if self._state is None:
            self._dev = dev
            self._state = torch.zeros(4, dtype=torch.int32, device=dev)
            self._entry = torch.zeros(3, self.capacity, dtype=torch.int32, device=dev)
        for name, t in rows.items():
            want = (self.capacity,) + tuple(t.shape[1:])
            cur = self._pool.get(name)
            if cur is None or cur.shape != want or cur.dtype != t.dtype:
                if self._calls:
                    raise RuntimeError("input layout changed inside a pending window: call flush() first")
                self._pool[name] = torch.empty(want, dtype=t.dtype, device=dev)

    @property
    def pending_calls(self) -> int:
        return len(self._calls)

    # -- register ----------------------------------------------------------------------------------------------
    @torch.no_grad()
    def register(self, features: Dict[str, torch.Tensor], mask: Optional[torch.Tensor], results, num_samples: int,
                 seed: int, clip0: int) -> None:
        """results = (mean_logits (B, C), uncertainty (B,), labels (B,) int32, packed (B, C + 1)) as mc_forward returns
        them; they are patched in place by flush().  Asynchronous: no host round trip."""
        mean, unc, labels, packed = results
        dev = mean.device
        B, Cn = mean.shape
        if B == 0 or Cn < 2:
            return
        if self._params is not None and self._params != (int(num_samples), int(seed)):
            self.flush()
        rows = {"audio": features["audio"], "video": features["video"]}
        for k, v in rows.items():
            if not v.is_contiguous():
                rows[k] = v.contiguous()
            if (rows[k][0].numel() * rows[k].element_size()) % 4 != 0:
                raise ValueError(f"features['{k}']: rows must be a multiple of 4 bytes")
        if mask is not None:
            m = torch.as_tensor(mask)
            if m.device != dev or m.dtype != torch.float32 or not m.is_contiguous():
                m = m.to(device=dev, dtype=torch.float32).contiguous()
            rows["mask"] = m
        self._ensure(dev, rows)
        self._params = (int(num_samples), int(seed))
        tag = self._next_tag
        self._next_tag = (self._next_tag + 1) & 0x7FFFFFFF
        lib = _lib.load()
        with torch.cuda.device(dev):
            s = _lib.stream_ptr(dev)
            ent = self._entry
            _lib.check(lib.avhot_flag_near_ties(_lib.ptr(mean), B, Cn, self.margin, tag, int(clip0) & (2**64 - 1),
                                                _lib.ptr(self._state), _lib.ptr(ent[0]), _lib.ptr(ent[1]),
                                                _lib.ptr(ent[2]), self.capacity, s), "flag_near_ties")
            for name, t in rows.items():
                _lib.check(lib.avhot_gather_flagged(_lib.ptr(t), t[0].numel() * t.element_size(), B,
                                                    _lib.ptr(self._state), _lib.ptr(ent[1]), _lib.ptr(self._pool[name]),
                                                    self.capacity, s), "gather_flagged")
        self._calls.append((tag, mean, unc, labels, packed, mask is not None))
        self.clips_total += B

    # -- flush -------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def flush(self, raise_on_overflow: bool = True) -> int:
        """Re-evaluate the pending near-tie clips through the fp32 kernels and patch the registered results.  One host
        round trip (the entry count).  Returns the number of clips re-evaluated (-1: more clips were flagged than the
        pool holds and raise_on_overflow is False - nothing was patched, the caller must redo the window in fp32)."""
        if not self._calls:
            return 0
        dev = self._dev
        st = self._state.cpu()
        n, overflow = int(st[0].item()), int(st[1].item())
        calls, (num_samples, seed) = self._calls, self._params
        self._calls, self._params = [], None
        self._state.zero_()
        if (overflow or n > self.capacity) and not raise_on_overflow:
            return -1
        if overflow or n > self.capacity:
            raise RuntimeError(f"NearTieResolver: {n} near-tie clips pending, capacity {self.capacity}: flush more often "
                               "or raise the capacity (the affected results are NOT label-exact)")
        if n == 0:
            return 0
        model = self.model
        clip_ids = self._entry[2, :n].contiguous()
        feats = {"audio": self._pool["audio"][:n], "video": self._pool["video"][:n]}
        has_mask = calls[0][5]
        if any(c[5] != has_mask for c in calls):
            raise RuntimeError("NearTieResolver: calls with and without a modality mask in one window")
        mask = self._pool["mask"][:n] if has_mask else None
        prev = model.precision
        model.set_precision("fp32")
        try:
            logits = model._run(feats, mask, num_samples, True, seed, 0, 0, clip_ids=clip_ids)
        finally:
            model.set_precision(prev)
        from .uncertainty import mc_reduce
        _, _, r_labels, r_packed = mc_reduce(logits)
        lib = _lib.load()
        Cn = r_packed.shape[1] - 1
        with torch.cuda.device(dev):
            for tag, mean, unc, labels, packed, _ in calls:
                _lib.check(lib.avhot_scatter_resolved(_lib.ptr(self._entry[0]), _lib.ptr(self._entry[1]), n, tag,
                                                      _lib.ptr(r_packed), _lib.ptr(r_labels), _lib.ptr(mean),
                                                      _lib.ptr(unc), _lib.ptr(labels), _lib.ptr(packed), Cn,
                                                      _lib.stream_ptr(dev)), "scatter_resolved")
        self.flagged_total += n
        return n
