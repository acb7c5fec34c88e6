"""Drop-in for src/attention.py: CrossModalAttention (the only attention class the fusion path uses,
src/fusion.py:14).  Same signature, parameter names and return values; forward runs libavhot.so."""
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
