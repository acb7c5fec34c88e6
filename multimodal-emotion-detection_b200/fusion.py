"""Drop-in for the cross-modal path of src/fusion.py: HybridFusion and build_fusion_model, plus the live
baseline head of src/train.py:81-85 as ConcatHead.  Same signatures, parameter names and return values."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib, _rng
from ._handles import Handle, Workspace
from .attention import CrossModalAttention


_ws = Workspace()


def _z_args(zs, rows_per_clip):
    M = len(zs)
    arr = (C.c_void_p * M)(*[z.data_ptr() for z in zs])
    rpc = (C.c_int * M)(*rows_per_clip)
    return arr, rpc


class HybridFusion(nn.Module):
    """src/fusion.py:404-530."""

    def __init__(self, modality_dims: Dict[str, int], hidden_dim: int = 256, num_classes: int = 11,
                 num_heads: int = 4, dropout: float = 0.1):
        super().__init__()
        self.modality_names = list(modality_dims.keys())
        self.num_modalities = len(self.modality_names)
        self.hidden_dim = hidden_dim
        self.num_heads = num_heads
        self.num_classes = num_classes
        self.modality_dims = dict(modality_dims)
        self.proj_layers = nn.ModuleDict({n: nn.Linear(d, hidden_dim) for n, d in modality_dims.items()})
        self.attn_layers = nn.ModuleDict({
            n: CrossModalAttention(query_dim=hidden_dim, key_dim=hidden_dim, hidden_dim=hidden_dim,
                                   num_heads=num_heads, dropout=dropout) for n in self.modality_names})
        for i, n in enumerate(self.modality_names):
            self.attn_layers[n].site_index = i
        self.pre_ln = nn.LayerNorm(hidden_dim)
        self.post_ln = nn.LayerNorm(hidden_dim)
        gate_hidden = max(32, hidden_dim // 2)
        self.gate_hidden = gate_hidden
        self.fusion_gate = nn.Sequential(nn.Linear(hidden_dim, gate_hidden), nn.ReLU(), nn.Dropout(dropout),
                                         nn.Linear(gate_hidden, 1))
        self.classifier = nn.Linear(hidden_dim, num_classes)
        self.precision = _lib.PREC_FP32
        self._handle = Handle("avhot_fusion_destroy")

    def _params(self):
        return [p for p in self.parameters()]

    def _get_handle(self, dev):
        def create():
            keep = []
            d = _lib.FusionDesc()
            d.num_modalities, d.hidden_dim, d.num_heads = self.num_modalities, self.hidden_dim, self.num_heads
            d.num_classes, d.gate_hidden = self.num_classes, self.gate_hidden
            if self.num_modalities > _lib.MAX_MODALITIES:
                raise ValueError(f"at most {_lib.MAX_MODALITIES} modalities")

            def dp(t):
                t = _lib.f32c(t.detach())
                _lib.require_cuda(t)
                keep.append(t)
                return t.data_ptr()

            for i, n in enumerate(self.modality_names):
                d.modality_dim[i] = self.modality_dims[n]
                d.proj_w[i], d.proj_b[i] = dp(self.proj_layers[n].weight), dp(self.proj_layers[n].bias)
                self.attn_layers[n].fill_desc(d.attn[i], keep)
            d.pre_ln_w, d.pre_ln_b = dp(self.pre_ln.weight), dp(self.pre_ln.bias)
            d.post_ln_w, d.post_ln_b = dp(self.post_ln.weight), dp(self.post_ln.bias)
            d.gate0_w, d.gate0_b = dp(self.fusion_gate[0].weight), dp(self.fusion_gate[0].bias)
            d.gate3_w, d.gate3_b = dp(self.fusion_gate[3].weight), dp(self.fusion_gate[3].bias)
            d.cls_w, d.cls_b = dp(self.classifier.weight), dp(self.classifier.bias)
            out = C.c_void_p()
            _lib.check(_lib.load().avhot_fusion_create(C.byref(d), None, C.byref(out), _lib.stream_ptr(dev)),
                       "HybridFusion")
            return out.value

        return self._handle.get(self._params(), create)

    def _drop_p(self) -> float:
        mods = [self.fusion_gate[2]] + [self.attn_layers[n].attn_dropout for n in self.modality_names]
        modes = {m.training for m in mods}
        if len(modes) != 1:
            raise ValueError("HybridFusion dropout modules must all be in the same mode")
        ps = {float(m.p) for m in mods}
        if len(ps) != 1:
            raise ValueError("HybridFusion dropout modules must share one p (one ctor argument, src/fusion.py:416)")
        return ps.pop() if modes.pop() else 0.0

    @torch.no_grad()
    def run(self, zs, rows_per_clip, mask, B: int, passes: int, p: float, seed: int, clip0: int, t0: int,
            want_info: bool = False):
        """logits (B, passes, C) for per-modality embeddings zs[m] of (B * rows_per_clip[m], dim_m)."""
        dev = zs[0].device
        R = B * passes
        logits = torch.empty(B, passes, self.num_classes, device=dev, dtype=torch.float32)
        fw = attn = hatt = None
        if want_info:
            M = self.num_modalities
            fw = torch.empty(R, M, device=dev, dtype=torch.float32)
            attn = torch.empty(M, R, self.num_heads, M, device=dev, dtype=torch.float32)
            hatt = torch.empty(R, M, self.hidden_dim, device=dev, dtype=torch.float32)
        if B > 0:
            arr, rpc = _z_args(zs, rows_per_clip)
            with torch.cuda.device(dev):
                h = self._get_handle(dev)
                lib = _lib.load()
                ws = _ws.get(lib.avhot_fusion_workspace_bytes(h, B, passes), dev)
                _lib.check(lib.avhot_fusion_forward(h, arr, rpc, _lib.ptr(mask), _lib.ptr(logits), _lib.ptr(fw),
                                                    _lib.ptr(attn), _lib.ptr(hatt), B, passes, float(p), seed, clip0,
                                                    t0, self.precision, _lib.ptr(ws), ws.numel(),
                                                    _lib.stream_ptr(dev)),
                           "HybridFusion.forward")
        return logits, fw, attn, hatt

    def forward(self, modality_features: Dict[str, torch.Tensor], modality_mask: Optional[torch.Tensor] = None,
                return_attention: bool = False):
        ref = modality_features[self.modality_names[0]]
        dev = _lib.require_cuda(ref)
        _lib.check_device(dev)
        B = ref.size(0)
        if modality_mask is None:                                                        # :468-469
            mask_f = torch.ones(B, self.num_modalities, device=dev, dtype=torch.float32)
        else:
            mask_f = modality_mask.to(device=dev)
            if mask_f.dtype != torch.float32:
                mask_f = mask_f.float()
            if mask_f.shape != (B, self.num_modalities):
                raise ValueError(f"modality_mask must be (B, {self.num_modalities})")
        zs = []
        for i, name in enumerate(self.modality_names):
            x = modality_features.get(name)
            if x is None:                                                                # :480-485 (mutates the mask)
                x = torch.zeros(B, self.modality_dims[name], device=dev, dtype=torch.float32)
                mask_f[:, i] = 0.0
            if x.dim() != 2 or x.shape[1] != self.modality_dims[name]:
                raise ValueError(f"features['{name}'] must be (B, {self.modality_dims[name]})")
            zs.append(_lib.f32c(x.to(dev)))
        mask_f = mask_f.contiguous()
        seed, t, clip0 = _rng.current()
        logits, fw, attn, hatt = self.run(zs, [1] * len(zs), mask_f, B, 1, self._drop_p(), seed, clip0, t,
                                          want_info=return_attention)
        logits = logits[:, 0, :]
        if return_attention:                                                             # :524-529
            return logits, {
                "fusion_weights": fw,
                "per_modality_attention": {n: attn[i].unsqueeze(2) for i, n in enumerate(self.modality_names)},
                "H_att": hatt,
            }
        return logits


class ConcatHead(nn.Sequential):
    """The live fusion_head of MultimodalFusionModule: Linear -> ReLU -> Linear on the concatenated encoder
    outputs (src/train.py:81-85, :153-164).  State-dict keys 0.weight, 0.bias, 2.weight, 2.bias."""

    def __init__(self, in_dim: int, hidden_dim: int, num_classes: int, num_modalities: int = 2):
        super().__init__(nn.Linear(in_dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, num_classes))
        self.in_dim, self.hidden_dim, self.num_classes = in_dim, hidden_dim, num_classes
        self.num_modalities = num_modalities
        self._handle = Handle("avhot_fusion_destroy")

    def _get_handle(self, dev):
        ps = [self[0].weight, self[0].bias, self[2].weight, self[2].bias]

        def create():
            ts = [_lib.f32c(p.detach()) for p in ps]
            _lib.require_cuda(*ts)
            d = _lib.HeadDesc()
            d.in_dim, d.hidden_dim, d.num_classes = self.in_dim, self.hidden_dim, self.num_classes
            d.w0, d.b0, d.w2, d.b2 = [t.data_ptr() for t in ts]
            out = C.c_void_p()
            _lib.check(_lib.load().avhot_fusion_create(None, C.byref(d), C.byref(out), _lib.stream_ptr(dev)),
                       "ConcatHead")
            return out.value

        return self._handle.get(ps, create)

    @torch.no_grad()
    def run(self, zs, rows_per_clip, B: int, passes: int):
        dev = zs[0].device
        logits = torch.empty(B, passes, self.num_classes, device=dev, dtype=torch.float32)
        if B > 0:
            arr, rpc = _z_args(zs, rows_per_clip)
            with torch.cuda.device(dev):
                h = self._get_handle(dev)
                _lib.check(_lib.load().avhot_fusion_forward(h, arr, rpc, None, _lib.ptr(logits), None, None, None,
                                                            B, passes, 0.0, 0, 0, 0, _lib.PREC_FP32, None, 0,
                                                            _lib.stream_ptr(dev)),
                           "ConcatHead.forward")
        return logits

    def forward(self, fused: torch.Tensor) -> torch.Tensor:
        dev = _lib.require_cuda(fused)
        _lib.check_device(dev)
        fused = _lib.f32c(fused)
        B = fused.shape[0]
        if fused.dim() != 2 or fused.shape[1] != self.in_dim:
            raise ValueError(f"expected (B, {self.in_dim})")
        step = self.in_dim // self.num_modalities
        zs = [fused[:, i * step:(i + 1) * step].contiguous() for i in range(self.num_modalities)]
        return self.run(zs, [1] * len(zs), B, 1)[:, 0, :]


def build_fusion_model(fusion_type: str, modality_dims: Dict[str, int], num_classes: int, **kwargs) -> nn.Module:
    """src/fusion.py:659-702.  'hybrid' is the cross-modal path; 'early' / 'late' are valid in the reference and
    not built here."""
    if fusion_type == "hybrid":
        return HybridFusion(modality_dims=modality_dims, num_classes=num_classes, **kwargs)
    if fusion_type in ("early", "late"):
        raise ValueError(f"fusion type '{fusion_type}' exists in the reference but is not built in this library "
                         "(hot path = 'hybrid')")
    raise ValueError(f"Unknown fusion type: {fusion_type}")                              # :696
