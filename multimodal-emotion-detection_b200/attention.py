"""Drop-in for src/attention.py: CrossModalAttention (the only attention class the fusion path uses,
src/fusion.py:14), and - SURVEY.md 8(f) row 4 - TemporalAttention and PairwiseModalityAttention.  Same signatures,
parameter names and return values; forward runs libavhot.so."""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, _rng
from ._handles import Handle, Workspace

_ws = Workspace()


class CrossModalAttention(nn.Module):
    """src/attention.py:14-141.  mask: keys, True = INVALID; numeric masks are 1 = valid and are inverted
    (:57-78)."""

    def __init__(self, query_dim: int, key_dim: int, hidden_dim: int, num_heads: int = 4, dropout: float = 0.1):
        super().__init__()
        assert hidden_dim % num_heads == 0, "hidden_dim must be divisible by num_heads"   # :34
        self.query_dim, self.key_dim = query_dim, key_dim
        self.hidden_dim = hidden_dim
        self.num_heads = num_heads
        self.head_dim = hidden_dim // num_heads
        self.scale = 1.0 / math.sqrt(self.head_dim)
        self.q_in_ln = nn.LayerNorm(query_dim)
        self.k_in_ln = nn.LayerNorm(key_dim)
        self.v_in_ln = nn.LayerNorm(key_dim)
        self.q_proj = nn.Linear(query_dim, hidden_dim, bias=True)
        self.k_proj = nn.Linear(key_dim, hidden_dim, bias=True)
        self.v_proj = nn.Linear(key_dim, hidden_dim, bias=True)
        self.out_proj = nn.Linear(hidden_dim, hidden_dim, bias=True)
        self.attn_dropout = nn.Dropout(dropout)
        self.site_index = 0          # Philox site = SITE_ATTN + site_index (HybridFusion numbers its layers)
        self._handle = Handle("avhot_xattn_destroy")

    def _ensure_seq(self, x: torch.Tensor) -> torch.Tensor:
        return x.unsqueeze(1) if x.dim() == 2 else x

    def _normalize_mask(self, mask: torch.Tensor, B: int, Tk: int) -> torch.Tensor:
        invalid = mask if mask.dtype == torch.bool else (mask <= 0)
        if invalid.dim() == 1:
            invalid = invalid.view(B, 1).expand(B, Tk)
        elif invalid.dim() == 2:
            if invalid.size(1) == 1:
                invalid = invalid.expand(B, Tk)
            else:
                assert invalid.size(1) == Tk, f"Mask width {invalid.size(1)} != Tk {Tk}"   # :74
        else:
            raise ValueError(f"Mask must be [B] or [B,Tk], got {invalid.shape}")           # :76
        return invalid

    def _params(self):
        return [self.q_in_ln.weight, self.q_in_ln.bias, self.k_in_ln.weight, self.k_in_ln.bias,
                self.v_in_ln.weight, self.v_in_ln.bias, self.q_proj.weight, self.q_proj.bias,
                self.k_proj.weight, self.k_proj.bias, self.v_proj.weight, self.v_proj.bias,
                self.out_proj.weight, self.out_proj.bias]

    def fill_desc(self, d: "_lib.XattnDesc", keep: list) -> None:
        ps = [_lib.f32c(p.detach()) for p in self._params()]
        _lib.require_cuda(*ps)
        keep.extend(ps)
        d.query_dim, d.key_dim, d.hidden_dim, d.num_heads = self.query_dim, self.key_dim, self.hidden_dim, self.num_heads
        (d.q_ln_w, d.q_ln_b, d.k_ln_w, d.k_ln_b, d.v_ln_w, d.v_ln_b, d.q_w, d.q_b, d.k_w, d.k_b, d.v_w, d.v_b,
         d.o_w, d.o_b) = [p.data_ptr() for p in ps]

    def _get_handle(self, dev):
        def create():
            d = _lib.XattnDesc()
            keep = []
            self.fill_desc(d, keep)
            out = C.c_void_p()
            _lib.check(_lib.load().avhot_xattn_create(C.byref(d), C.byref(out), _lib.stream_ptr(dev)),
                       "CrossModalAttention")
            return out.value

        return self._handle.get(self._params(), create)

    @torch.no_grad()
    def forward(self, query: torch.Tensor, key: torch.Tensor, value: torch.Tensor,
                mask: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        dev = _lib.require_cuda(query, key, value)
        _lib.check_device(dev)
        q = _lib.f32c(self._ensure_seq(query))
        k = _lib.f32c(self._ensure_seq(key))
        v = _lib.f32c(self._ensure_seq(value))
        B, Tq, Dq = q.shape
        _, Tk, Dk = k.shape
        if Dq != self.query_dim or Dk != self.key_dim or v.shape != k.shape:
            raise ValueError(f"shapes do not match the module: query {tuple(q.shape)}, key {tuple(k.shape)}, "
                             f"value {tuple(v.shape)}")
        invalid = None
        if mask is not None:
            invalid = self._normalize_mask(mask.to(dev), B, Tk).to(torch.uint8).contiguous()
        out = torch.empty(B, Tq, self.hidden_dim, device=dev, dtype=torch.float32)
        attn = torch.empty(B, self.num_heads, Tq, Tk, device=dev, dtype=torch.float32)
        if B > 0:
            p = float(self.attn_dropout.p) if self.attn_dropout.training else 0.0
            seed, t, clip0 = _rng.current()
            lib = _lib.load()
            with torch.cuda.device(dev):
                h = self._get_handle(dev)
                ws = _ws.get(lib.avhot_xattn_workspace_bytes(h, B, Tq, Tk), dev)
                _lib.check(lib.avhot_xattn_forward(h, _lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(invalid),
                                                   _lib.ptr(out), _lib.ptr(attn), B, Tq, Tk, p, seed, clip0, t,
                                                   self.site_index, _lib.ptr(ws), ws.numel(),
                                                   _lib.stream_ptr(dev)), "CrossModalAttention.forward")
        out = out.squeeze(1) if Tq == 1 else out                                         # :140
        return out, attn


class TemporalAttention(nn.Module):
    """src/attention.py:143-274: multi-head self-attention over the time steps of one sequence.  The reference's
    forward never applies its two nn.Dropout modules (it tests `hasattr(self, "attn_drop")` / `"proj_drop"`, :227 and
    :240, while the attributes are attn_dropout / proj_dropout), so the module is deterministic in every mode; the
    modules are kept for state-dict / attribute parity."""

    def __init__(self, feature_dim: int, hidden_dim: int = 256, num_heads: int = 4, dropout: float = 0.1):
        super().__init__()
        self.feature_dim = feature_dim
        self.hidden_dim = hidden_dim
        self.num_heads = num_heads
        self.head_dim = hidden_dim // num_heads
        self.q_proj = nn.Linear(feature_dim, hidden_dim)
        self.k_proj = nn.Linear(feature_dim, hidden_dim)
        self.v_proj = nn.Linear(feature_dim, hidden_dim)
        self.out_proj = nn.Linear(hidden_dim, hidden_dim)
        self.attn_dropout = nn.Dropout(dropout)
        self.proj_dropout = nn.Dropout(dropout)
        self.scale = (self.head_dim) ** -0.5
        self._handle = Handle("avhot_tattn_destroy")

    def _params(self):
        return [self.q_proj.weight, self.q_proj.bias, self.k_proj.weight, self.k_proj.bias, self.v_proj.weight,
                self.v_proj.bias, self.out_proj.weight, self.out_proj.bias]

    def _get_handle(self, dev):
        src = self._params()

        def create():
            if self.head_dim * self.num_heads != self.hidden_dim:
                raise ValueError("hidden_dim must be divisible by num_heads")   # the reference's .view fails (:214)
            ts = [_lib.f32c(t.detach()) for t in src]
            _lib.require_cuda(*ts)
            d = _lib.TattnDesc()
            d.feature_dim, d.hidden_dim, d.num_heads = self.feature_dim, self.hidden_dim, self.num_heads
            d.q_w, d.q_b, d.k_w, d.k_b, d.v_w, d.v_b, d.o_w, d.o_b = [t.data_ptr() for t in ts]
            out = C.c_void_p()
            _lib.check(_lib.load().avhot_tattn_create(C.byref(d), C.byref(out), _lib.stream_ptr(dev)), "TemporalAttention")
            torch.cuda.current_stream(dev).synchronize()       # the handle copies from `ts` on the stream
            return out.value

        return self._handle.get(src, create)

    @torch.no_grad()
    def forward(self, sequence: torch.Tensor, mask: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """sequence (B, S, feature_dim), mask (B, S) 1 = valid -> (attended (B, S, hidden), weights (B, heads, S, S))."""
        dev = _lib.require_cuda(sequence)
        _lib.check_device(dev)
        if sequence.dim() != 3 or sequence.shape[2] != self.feature_dim:
            raise ValueError(f"sequence must be (B, S, {self.feature_dim}), got {tuple(sequence.shape)}")
        x = _lib.f32c(sequence)
        B, S, _ = x.shape
        if S < 1:
            raise ValueError("empty sequence")
        valid = None
        if mask is not None:
            valid = mask.to(device=dev, dtype=torch.bool).to(torch.uint8).contiguous()     # :220, :233
            if valid.shape != (B, S):
                raise ValueError("mask must be (B, S)")
        attended = torch.empty(B, S, self.hidden_dim, device=dev, dtype=torch.float32)
        weights = torch.empty(B, self.num_heads, S, S, device=dev, dtype=torch.float32)
        if B > 0:
            lib = _lib.load()
            with torch.cuda.device(dev):
                h = self._get_handle(dev)
                ws = _ws.get(lib.avhot_tattn_workspace_bytes(h, B, S), dev)
                _lib.check(lib.avhot_tattn_forward(h, _lib.ptr(x), _lib.ptr(valid), _lib.ptr(attended), _lib.ptr(weights),
                                                   B, S, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)),
                           "TemporalAttention.forward")
        return attended, weights

    @torch.no_grad()
    def pool_sequence(self, sequence: torch.Tensor, attention_weights: torch.Tensor) -> torch.Tensor:
        """src/attention.py:245-274: (B, S, D), (B, heads, S, S) -> (B, D)."""
        dev = _lib.require_cuda(sequence, attention_weights)
        _lib.check_device(dev)
        x, w = _lib.f32c(sequence), _lib.f32c(attention_weights)
        B, S, D = x.shape
        if w.dim() != 4 or w.shape[0] != B or w.shape[2] != S or w.shape[3] != S:
            raise ValueError(f"attention_weights must be (B, heads, {S}, {S}), got {tuple(w.shape)}")
        pooled = torch.empty(B, D, device=dev, dtype=torch.float32)
        if B > 0:
            with torch.cuda.device(dev):
                _lib.check(_lib.load().avhot_tattn_pool(_lib.ptr(x), _lib.ptr(w), _lib.ptr(pooled), B, S, D, w.shape[1],
                                                        _lib.stream_ptr(dev)), "TemporalAttention.pool_sequence")
        return pooled


class PairwiseModalityAttention(nn.Module):
    """src/attention.py:421-548 (the file defines the class twice; this is the second definition, the one the name is
    bound to): one CrossModalAttention per direction of every modality pair, a per-modality self projection, and
    out = LayerNorm(self_proj(x) + msg_dropout(sum of incoming messages)), zeroed where the modality is missing.
    Dropout sites: direction k of `cross_attn` (ModuleDict order) draws SITE_ATTN + k, msg_dropout for target
    modality i draws SITE_PAIR_MSG + i.  At most 4 modalities."""

    def __init__(self, modality_dims: dict, hidden_dim: int = 256, num_heads: int = 4, dropout: float = 0.1):
        super().__init__()
        self.modality_names = list(modality_dims.keys())
        self.num_modalities = len(self.modality_names)
        self.hidden_dim = hidden_dim
        if self.num_modalities > _lib.MAX_MODALITIES:
            raise ValueError(f"at most {_lib.MAX_MODALITIES} modalities are built")
        self.cross_attn = nn.ModuleDict()
        for i in range(self.num_modalities):
            for j in range(i + 1, self.num_modalities):
                a, b = self.modality_names[i], self.modality_names[j]
                da, db = modality_dims[a], modality_dims[b]
                self.cross_attn[f"{a}_to_{b}"] = CrossModalAttention(da, db, hidden_dim, num_heads, dropout)
                self.cross_attn[f"{b}_to_{a}"] = CrossModalAttention(db, da, hidden_dim, num_heads, dropout)
        for k, key in enumerate(self.cross_attn.keys()):
            self.cross_attn[key].site_index = k
        self.self_proj = nn.ModuleDict({m: nn.Linear(modality_dims[m], hidden_dim) for m in self.modality_names})
        self.out_ln = nn.LayerNorm(hidden_dim)
        self.msg_dropout = nn.Dropout(dropout)

    @torch.no_grad()
    def forward(self, modality_features: dict, modality_mask: Optional[torch.Tensor] = None) -> Tuple[dict, dict]:
        names = self.modality_names
        first = next(iter(modality_features.values()))
        dev = _lib.require_cuda(*modality_features.values())
        _lib.check_device(dev)
        B = first.size(0)
        if modality_mask is not None:
            assert modality_mask.shape == (B, self.num_modalities), f"modality_mask must be (B,{self.num_modalities})"
            mm = modality_mask.to(dev)
            avail = {m: ((mm[:, i] > 0) if mm.dtype != torch.bool else mm[:, i]).contiguous() for i, m in enumerate(names)}
            # the reference's `b_ok.any()` / `avail.all()` tests read the mask on the host (:497, :509, :539); one copy here
            any_avail = {m: bool(v) for m, v in zip(names, torch.stack([avail[m].any() for m in names]).tolist())} if B > 0 \
                else {m: False for m in names}
        else:
            avail = {m: None for m in names}
            any_avail = {m: True for m in names}
        feats = {m: _lib.f32c(modality_features[m]) for m in names}
        messages = {m: [] for m in names}
        attention_maps: dict = {}
        for i in range(self.num_modalities):
            for j in range(i + 1, self.num_modalities):
                a, b = names[i], names[j]
                for tgt, src in ((a, b), (b, a)):
                    if any_avail[src]:                                                   # :497, :509
                        inv = None if avail[src] is None else ~avail[src]
                        out, att = self.cross_attn[f"{tgt}_to_{src}"](feats[tgt], feats[src], feats[src], mask=inv)
                        messages[tgt].append(out)
                        attention_maps[f"{tgt}_to_{src}"] = att
        p = float(self.msg_dropout.p) if self.msg_dropout.training else 0.0
        seed, t, clip0 = _rng.current()
        lib = _lib.load()
        attended = {}
        for i, m in enumerate(names):
            x = feats[m]
            self_feat = torch.empty(B, self.hidden_dim, device=dev, dtype=torch.float32)
            agg = torch.empty(B, self.hidden_dim, device=dev, dtype=torch.float32)
            if B > 0:
                w, bias = _lib.f32c(self.self_proj[m].weight.detach()), _lib.f32c(self.self_proj[m].bias.detach())
                ln_w, ln_b = _lib.f32c(self.out_ln.weight.detach()), _lib.f32c(self.out_ln.bias.detach())
                _lib.require_cuda(w, bias, ln_w, ln_b)
                av = None if avail[m] is None else avail[m].to(torch.uint8)
                msgs = messages[m]
                arr = (C.c_void_p * max(len(msgs), 1))(*[t.data_ptr() for t in msgs])
                with torch.cuda.device(dev):
                    _lib.check(lib.avhot_gemm(_lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(self_feat), B,
                                              self.hidden_dim, x.shape[1], 0, _lib.PREC_FP32, _lib.stream_ptr(dev)),
                               "PairwiseModalityAttention.self_proj")
                    _lib.check(lib.avhot_pair_combine(_lib.ptr(self_feat), arr, len(msgs), _lib.ptr(av), _lib.ptr(ln_w),
                                                      _lib.ptr(ln_b), float(self.out_ln.eps), _lib.ptr(agg), B,
                                                      self.hidden_dim, p, seed, clip0, t, _lib.SITE_PAIR_MSG + i,
                                                      _lib.stream_ptr(dev)), "PairwiseModalityAttention")
            attended[m] = agg
        return attended, attention_maps
