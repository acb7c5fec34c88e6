"""Drop-in for src/fusion.py: HybridFusion (the cross-modal hot path), EarlyFusion, LateFusion and
build_fusion_model, plus the live baseline head of src/train.py:81-85 as ConcatHead.  Same signatures, parameter
names and return values."""
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
        self._handle_folded = Handle("avhot_fusion_destroy")

    def _params(self):
        return [p for p in self.parameters()]

    def _get_handle(self, dev, folded=None):
        """folded = (index, in_dim, weight (hidden, in_dim), bias (hidden), source tensors): a second handle whose
        proj_layers[index] is replaced by a projection the caller has folded an upstream linear map into."""
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
                if folded is not None and folded[0] == i:
                    d.modality_dim[i] = folded[1]
                    d.proj_w[i], d.proj_b[i] = dp(folded[2]), dp(folded[3])
                self.attn_layers[n].fill_desc(d.attn[i], keep)
            d.pre_ln_w, d.pre_ln_b = dp(self.pre_ln.weight), dp(self.pre_ln.bias)
            d.post_ln_w, d.post_ln_b = dp(self.post_ln.weight), dp(self.post_ln.bias)
            d.gate0_w, d.gate0_b = dp(self.fusion_gate[0].weight), dp(self.fusion_gate[0].bias)
            d.gate3_w, d.gate3_b = dp(self.fusion_gate[3].weight), dp(self.fusion_gate[3].bias)
            d.cls_w, d.cls_b = dp(self.classifier.weight), dp(self.classifier.bias)
            out = C.c_void_p()
            _lib.check(_lib.load().avhot_fusion_create(C.byref(d), None, C.byref(out), _lib.stream_ptr(dev)),
                       "HybridFusion")
            if folded is not None:
                torch.cuda.current_stream(dev).synchronize()   # the handle packs from the folded temporaries on the stream
            return out.value

        if folded is not None:
            return self._handle_folded.get(self._params() + list(folded[4]), create)
        return self._handle.get(self._params(), create)

    def _fold_video_projection(self, index: int, lin: nn.Linear, dev):
        """proj_layers[index] o lin as one Linear(lin.in_features -> hidden): W' = W_p W_l, b' = W_p b_l + b_p, computed
        with the library's fp32 GEMM (two linear maps with no non-linearity between them: FrameEncoder.projection[1],
        src/encoders.py:249, feeding proj_layers.video, src/fusion.py:486)."""
        name = self.modality_names[index]
        pl = self.proj_layers[name]
        w_p, b_p = _lib.f32c(pl.weight.detach()), _lib.f32c(pl.bias.detach())              # (hidden, out), (hidden)
        w_l, b_l = _lib.f32c(lin.weight.detach()), _lib.f32c(lin.bias.detach())            # (out, in), (out)
        _lib.require_cuda(w_p, b_p, w_l, b_l)
        w_lt = w_l.t().contiguous()                                                        # (in, out): layout copy only
        hidden, k, n_in = w_p.shape[0], w_p.shape[1], w_l.shape[1]
        w = torch.empty(hidden, n_in, device=dev, dtype=torch.float32)
        b = torch.empty(1, hidden, device=dev, dtype=torch.float32)
        lib = _lib.load()
        with torch.cuda.device(dev):
            _lib.check(lib.avhot_gemm(_lib.ptr(w_p), _lib.ptr(w_lt), None, _lib.ptr(w), hidden, n_in, k, 0, _lib.PREC_FP32,
                                      _lib.stream_ptr(dev)), "HybridFusion fold (weight)")
            _lib.check(lib.avhot_gemm(_lib.ptr(b_l.view(1, -1)), _lib.ptr(w_p), _lib.ptr(b_p), _lib.ptr(b), 1, hidden, k, 0,
                                      _lib.PREC_FP32, _lib.stream_ptr(dev)), "HybridFusion fold (bias)")
        return w, b.view(-1)

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
            want_info: bool = False, folded_from=None, clip_ids=None):
        """logits (B, passes, C) for per-modality embeddings zs[m] of (B * rows_per_clip[m], dim_m).
        clip_ids: optional (B,) int32/uint32-valued CUDA tensor of global clip indices (instead of clip0 + b): the fp32
        re-evaluation of scattered near-tie clips (exact.NearTieResolver).
        folded_from = (index, nn.Linear): zs[index] is the INPUT of that Linear (whose output is modality `index`'s
        embedding); the Linear is folded into proj_layers[index] of a second handle."""
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
                if folded_from is not None:
                    idx, lin = folded_from
                    src = [lin.weight, lin.bias]
                    w = b = None
                    if not self._handle_folded.is_current(self._params() + src):     # (re)fold only when a weight changed
                        w, b = self._fold_video_projection(idx, lin, dev)
                    h = self._get_handle(dev, folded=(idx, lin.in_features, w, b, src))
                else:
                    h = self._get_handle(dev)
                lib = _lib.load()
                ws = _ws.get(lib.avhot_fusion_workspace_bytes(h, B, passes), dev)
                if clip_ids is not None:
                    if want_info:
                        raise ValueError("clip_ids and want_info are mutually exclusive")
                    _lib.check(lib.avhot_fusion_forward_ids(h, arr, rpc, _lib.ptr(mask), _lib.ptr(logits), B, passes,
                                                            float(p), seed, _lib.ptr(clip_ids), t0, self.precision,
                                                            _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)),
                               "HybridFusion.forward")
                else:
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


class _HeadBase(nn.Module):
    """Shared run()/mask plumbing of EarlyFusion / LateFusion (avhot_head_forward)."""

    def _drop_p(self) -> float:
        mods = [m for m in self.modules() if isinstance(m, nn.Dropout)]
        modes = {m.training for m in mods}
        if len(modes) != 1:
            raise ValueError(f"{type(self).__name__} dropout modules must all be in the same mode")
        ps = {float(m.p) for m in mods}
        if len(ps) != 1:
            raise ValueError(f"{type(self).__name__} dropout modules must share one p (one ctor argument)")
        return ps.pop() if modes.pop() else 0.0

    @torch.no_grad()
    def run(self, zs, rows_per_clip, mask, B: int, passes: int, p: float, seed: int, clip0: int, t0: int,
            present: Optional[int] = None, want_per_modality: bool = False):
        """logits (B, passes, C) [and per-modality logits (M, B, passes, C)] for embeddings zs[m] of
        (B * rows_per_clip[m], dim_m); zs[m] may be None for a modality that was not given (LateFusion)."""
        dev = next(z for z in zs if z is not None).device
        M = len(zs)
        if present is None:
            present = sum(1 << m for m in range(M) if zs[m] is not None)
        logits = torch.empty(B, passes, self.num_classes, device=dev, dtype=torch.float32)
        per = torch.empty(M, B, passes, self.num_classes, device=dev, dtype=torch.float32) if want_per_modality else None
        if B > 0:
            arr = (C.c_void_p * M)(*[None if z is None else z.data_ptr() for z in zs])
            rpc = (C.c_int * M)(*rows_per_clip)
            with torch.cuda.device(dev):
                h = self._get_handle(dev)
                _lib.check(_lib.load().avhot_head_forward(h, arr, rpc, present, _lib.ptr(mask), _lib.ptr(logits),
                                                          _lib.ptr(per), B, passes, float(p), seed, clip0, t0,
                                                          _lib.stream_ptr(dev)),
                           f"{type(self).__name__}.forward")
        return logits, per

    def _ref_and_mask(self, modality_features, modality_mask):
        ref = None
        for name in self.modality_names:                                                 # src/fusion.py:93-99
            if name in modality_features:
                ref = modality_features[name]
                break
        if ref is None:
            raise ValueError("No expected modalities were provided.")
        if ref.dim() != 2:
            raise ValueError(f"Expected rank-2 modality tensors [B, D], got {tuple(ref.shape)}.")
        dev = _lib.require_cuda(ref)
        _lib.check_device(dev)
        B = ref.size(0)
        mask_f = None
        if modality_mask is not None:                                                    # :103-106
            if modality_mask.dim() != 2 or modality_mask.size(0) != B or \
                    modality_mask.size(1) != len(self.modality_names):
                raise ValueError(f"modality_mask must be [B, {len(self.modality_names)}]")
            mask_f = _lib.f32c(modality_mask.to(device=dev))
        return dev, B, mask_f


class EarlyFusion(_HeadBase):
    """src/fusion.py:19-138: concat(mask_m * z_m) -> Linear-LayerNorm-ReLU-Dropout x2 -> Linear (state-dict keys
    mlp.0, mlp.1, mlp.4, mlp.5, mlp.8)."""

    def __init__(self, modality_dims: Dict[str, int], hidden_dim: int = 256, num_classes: int = 11,
                 dropout: float = 0.1, num_heads: int = 4, **kwargs):
        super().__init__()
        self._num_heads = num_heads
        self.modality_names = list(modality_dims.keys())
        self.modality_dims = dict(modality_dims)
        self.hidden_dim, self.num_classes = hidden_dim, num_classes
        concat_dim = sum(modality_dims.values())
        self.mlp = nn.Sequential(                                                        # :60-72
            nn.Linear(concat_dim, hidden_dim), nn.LayerNorm(hidden_dim), nn.ReLU(inplace=True), nn.Dropout(dropout),
            nn.Linear(hidden_dim, hidden_dim), nn.LayerNorm(hidden_dim), nn.ReLU(inplace=True), nn.Dropout(dropout),
            nn.Linear(hidden_dim, num_classes))
        self._handle = Handle("avhot_head_destroy")

    def _get_handle(self, dev):
        ps = [p for p in self.parameters()]

        def create():
            if len(self.modality_names) > _lib.MAX_MODALITIES:
                raise ValueError(f"at most {_lib.MAX_MODALITIES} modalities")
            ts = [_lib.f32c(p.detach()) for p in ps]
            _lib.require_cuda(*ts)
            d = _lib.EarlyDesc()
            d.num_modalities, d.hidden_dim, d.num_classes = len(self.modality_names), self.hidden_dim, self.num_classes
            for i, n in enumerate(self.modality_names):
                d.modality_dim[i] = self.modality_dims[n]
            (d.w0, d.b0, d.ln1_w, d.ln1_b, d.w4, d.b4, d.ln5_w, d.ln5_b, d.w8, d.b8) = [t.data_ptr() for t in ts]
            out = C.c_void_p()
            _lib.check(_lib.load().avhot_early_create(C.byref(d), C.byref(out), _lib.stream_ptr(dev)), "EarlyFusion")
            torch.cuda.current_stream(dev).synchronize()       # the handle copies from `ts` on the stream
            return out.value

        return self._handle.get(ps, create)

    def forward(self, modality_features: Dict[str, torch.Tensor],
                modality_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        dev, B, mask_f = self._ref_and_mask(modality_features, modality_mask)
        zs = []
        for name in self.modality_names:
            Dm = self.modality_dims[name]
            x = modality_features.get(name)
            if x is None:                                                                # :120-121
                x = torch.zeros(B, Dm, device=dev, dtype=torch.float32)
            elif x.dim() != 2 or x.size(0) != B or x.size(1) != Dm:
                raise ValueError(f"Modality '{name}' must be [B, {Dm}], got {tuple(x.shape)}")
            zs.append(_lib.f32c(x.to(dev)))
        seed, t, clip0 = _rng.current()
        logits, _ = self.run(zs, [1] * len(zs), mask_f, B, 1, self._drop_p(), seed, clip0, t)
        return logits[:, 0, :]


class LateFusion(_HeadBase):
    """src/fusion.py:141-248: per-modality Linear-ReLU-Dropout-Linear classifiers; softmax(fusion_logits) times the
    mask, renormalised; returns (fused_logits, {name: per-modality logits})."""

    def __init__(self, modality_dims: Dict[str, int], hidden_dim: int = 256, num_classes: int = 11,
                 dropout: float = 0.1, num_heads: int = 4, **kwargs):
        super().__init__()
        self._num_heads = num_heads
        self.modality_names = list(modality_dims.keys())
        self.num_modalities = len(self.modality_names)
        self.modality_dims = dict(modality_dims)
        self.hidden_dim, self.num_classes = hidden_dim, num_classes
        self.classifiers = nn.ModuleDict({                                               # :180-188
            name: nn.Sequential(nn.Linear(dim, hidden_dim), nn.ReLU(), nn.Dropout(dropout),
                                nn.Linear(hidden_dim, num_classes))
            for name, dim in modality_dims.items()})
        self.fusion_logits = nn.Parameter(torch.zeros(self.num_modalities))              # :189
        self._handle = Handle("avhot_head_destroy")

    def _get_handle(self, dev):
        ps = [p for p in self.parameters()]

        def create():
            if self.num_modalities > _lib.MAX_MODALITIES:
                raise ValueError(f"at most {_lib.MAX_MODALITIES} modalities")
            keep = []

            def dp(t):
                t = _lib.f32c(t.detach())
                _lib.require_cuda(t)
                keep.append(t)
                return t.data_ptr()

            d = _lib.LateDesc()
            d.num_modalities, d.hidden_dim, d.num_classes = self.num_modalities, self.hidden_dim, self.num_classes
            for i, n in enumerate(self.modality_names):
                seq = self.classifiers[n]
                d.modality_dim[i] = self.modality_dims[n]
                d.w0[i], d.b0[i], d.w3[i], d.b3[i] = dp(seq[0].weight), dp(seq[0].bias), dp(seq[3].weight), dp(seq[3].bias)
            d.fusion_logits = dp(self.fusion_logits)
            out = C.c_void_p()
            _lib.check(_lib.load().avhot_late_create(C.byref(d), C.byref(out), _lib.stream_ptr(dev)), "LateFusion")
            torch.cuda.current_stream(dev).synchronize()
            return out.value

        return self._handle.get(ps, create)

    def forward(self, modality_features: Dict[str, torch.Tensor], modality_mask: Optional[torch.Tensor] = None):
        dev, B, mask_f = self._ref_and_mask(modality_features, modality_mask)
        if mask_f is not None and mask_f.shape != (B, self.num_modalities):
            raise ValueError(f"modality_mask must be [B, {self.num_modalities}]")
        zs = []
        for name in self.modality_names:
            x = modality_features.get(name)
            if x is not None:
                if x.dim() != 2 or x.size(0) != B or x.size(1) != self.modality_dims[name]:
                    raise ValueError(f"Modality '{name}' must be [B, {self.modality_dims[name]}], got {tuple(x.shape)}")
                x = _lib.f32c(x.to(dev))
            zs.append(x)                                                                 # None: zero logits (:222-225)
        seed, t, clip0 = _rng.current()
        logits, per = self.run(zs, [1] * len(zs), mask_f, B, 1, self._drop_p(), seed, clip0, t,
                               want_per_modality=True)
        return logits[:, 0, :], {n: per[i, :, 0, :] for i, n in enumerate(self.modality_names)}


class LateFusionWithUncertainty(nn.Module):
    """src/fusion.py:589-655 (build_fusion_model('uncertainty')): per modality a classifier and a scalar-uncertainty
    head (softplus), fused by inverse uncertainty (UncertaintyWeightedFusion).  forward(encoded_features,
    modality_mask) -> (fused_logits, {"per_modality_logits": (B, M, C), "fusion_weights": (B, M)})."""

    def __init__(self, modality_dims: Dict[str, int], num_classes: int, hidden_dim: int = 0, num_heads: int = 0,
                 dropout: float = 0.0, epsilon: float = 1e-6):
        super().__init__()
        from .uncertainty import UncertaintyWeightedFusion
        self.modalities = list(modality_dims.keys())
        self.modality_dims = modality_dims
        self.num_classes = num_classes
        self.hidden_dim = hidden_dim if hidden_dim and hidden_dim > 0 else 0
        self.classifiers = nn.ModuleDict()
        self.uncert_heads = nn.ModuleDict()
        for m, d in modality_dims.items():                                               # :611-627
            if self.hidden_dim:
                self.classifiers[m] = nn.Sequential(nn.Dropout(dropout), nn.Linear(d, hidden_dim), nn.ReLU(inplace=True),
                                                    nn.Dropout(dropout), nn.Linear(hidden_dim, num_classes))
                self.uncert_heads[m] = nn.Sequential(nn.Dropout(dropout), nn.Linear(d, hidden_dim), nn.ReLU(inplace=True),
                                                     nn.Dropout(dropout), nn.Linear(hidden_dim, 1))
            else:
                self.classifiers[m] = nn.Sequential(nn.Dropout(dropout), nn.Linear(d, num_classes))
                self.uncert_heads[m] = nn.Sequential(nn.Dropout(dropout), nn.Linear(d, 1))
        self.fuser = UncertaintyWeightedFusion(epsilon=epsilon)
        self._handle = Handle("avhot_head_destroy")

    _drop_p = _HeadBase._drop_p

    def _get_handle(self, dev):
        ps = [p for p in self.parameters()]

        def create():
            if len(self.modalities) > _lib.MAX_MODALITIES:
                raise ValueError(f"at most {_lib.MAX_MODALITIES} modalities")
            keep = []

            def dp(t):
                t = _lib.f32c(t.detach())
                _lib.require_cuda(t)
                keep.append(t)
                return t.data_ptr()

            d = _lib.UncDesc()
            d.num_modalities, d.hidden_dim, d.num_classes = len(self.modalities), self.hidden_dim, self.num_classes
            d.epsilon = float(self.fuser.epsilon)
            for i, n in enumerate(self.modalities):
                d.modality_dim[i] = self.modality_dims[n]
                c, u = self.classifiers[n], self.uncert_heads[n]
                if self.hidden_dim:
                    d.cls_w0[i], d.cls_b0[i], d.cls_w3[i], d.cls_b3[i] = dp(c[1].weight), dp(c[1].bias), dp(c[4].weight), dp(c[4].bias)
                    d.unc_w0[i], d.unc_b0[i], d.unc_w3[i], d.unc_b3[i] = dp(u[1].weight), dp(u[1].bias), dp(u[4].weight), dp(u[4].bias)
                else:
                    d.cls_w3[i], d.cls_b3[i] = dp(c[1].weight), dp(c[1].bias)
                    d.unc_w3[i], d.unc_b3[i] = dp(u[1].weight), dp(u[1].bias)
            out = C.c_void_p()
            _lib.check(_lib.load().avhot_unc_create(C.byref(d), C.byref(out), _lib.stream_ptr(dev)),
                       "LateFusionWithUncertainty")
            torch.cuda.current_stream(dev).synchronize()
            return out.value

        return self._handle.get(ps, create)

    @torch.no_grad()
    def run(self, zs, rows_per_clip, mask, B: int, passes: int, p: float, seed: int, clip0: int, t0: int):
        dev = mask.device
        M = len(zs)
        present = sum(1 << m for m in range(M) if zs[m] is not None)
        logits = torch.empty(B, passes, self.num_classes, device=dev, dtype=torch.float32)
        per = torch.empty(B, passes, M, self.num_classes, device=dev, dtype=torch.float32)
        fw = torch.empty(B, passes, M, device=dev, dtype=torch.float32)
        if B > 0:
            arr = (C.c_void_p * M)(*[None if z is None else z.data_ptr() for z in zs])
            rpc = (C.c_int * M)(*rows_per_clip)
            with torch.cuda.device(dev):
                h = self._get_handle(dev)
                _lib.check(_lib.load().avhot_unc_forward(h, arr, rpc, present, _lib.ptr(mask), _lib.ptr(logits),
                                                         _lib.ptr(per), _lib.ptr(fw), B, passes, float(p), seed, clip0,
                                                         t0, _lib.stream_ptr(dev)), "LateFusionWithUncertainty.forward")
        return logits, per, fw

    def forward(self, encoded_features: Dict[str, torch.Tensor], modality_mask: torch.Tensor):
        assert modality_mask is not None, "modality_mask (B,M) is required."           # :622
        dev = _lib.require_cuda(modality_mask, *encoded_features.values())
        _lib.check_device(dev)
        B = modality_mask.size(0)
        mask_f = _lib.f32c(modality_mask)
        if mask_f.shape != (B, len(self.modalities)):
            raise ValueError(f"modality_mask must be [B, {len(self.modalities)}]")
        zs = []
        for m in self.modalities:
            x = encoded_features.get(m)
            if x is not None:
                if x.size(0) != B:
                    raise ValueError(f"Batch mismatch for modality '{m}': {x.size(0)} vs {B}")   # :631
                if x.dim() != 2 or x.size(1) != self.modality_dims[m]:
                    raise ValueError(f"Modality '{m}' must be [B, {self.modality_dims[m]}], got {tuple(x.shape)}")
                x = _lib.f32c(x)
            zs.append(x)
        seed, t, clip0 = _rng.current()
        logits, per, fw = self.run(zs, [1] * len(zs), mask_f, B, 1, self._drop_p(), seed, clip0, t)
        return logits[:, 0, :], {"per_modality_logits": per[:, 0], "fusion_weights": fw[:, 0]}


class ConcatHead(nn.Sequential):
    """The live fusion_head of MultimodalFusionModule: Linear -> ReLU -> Linear on the concatenated encoder
    outputs (src/train.py:81-85, :153-164).  State-dict keys 0.weight, 0.bias, 2.weight, 2.bias."""

    def __init__(self, in_dim: int, hidden_dim: int, num_classes: int, num_modalities: int = 2):
        # the kernel (concat_head_kernel, csrc/fusion.cu) reads exactly two embeddings of in_dim / 2 columns each
        if num_modalities != 2 or in_dim % 2 != 0 or (in_dim // 2) % 4 != 0:
            raise ValueError(f"ConcatHead is built for two modalities of equal width (a multiple of 4): got "
                             f"num_modalities={num_modalities}, in_dim={in_dim}")
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
        if len(zs) != 2 or len(rows_per_clip) != 2:
            raise ValueError(f"ConcatHead.run takes exactly two modality embeddings, got {len(zs)}")
        for z in zs:
            if z.dim() != 2 or z.shape[1] != self.in_dim // 2:
                raise ValueError(f"ConcatHead.run: embeddings must be (rows, {self.in_dim // 2}), got {tuple(z.shape)}")
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
    """src/fusion.py:659-702.  'hybrid' is the cross-modal path; 'early' / 'late' / 'uncertainty'
    (LateFusionWithUncertainty) are SURVEY.md 8(f) rows."""
    if fusion_type in {"uncertainty"}:                                                   # :676-686
        return LateFusionWithUncertainty(modality_dims=modality_dims, num_classes=num_classes,
                                         hidden_dim=kwargs.get("hidden_dim", 0), num_heads=kwargs.get("num_heads", 0),
                                         dropout=kwargs.get("dropout", 0.0), epsilon=kwargs.get("epsilon", 1e-6))
    fusion_classes = {"early": EarlyFusion, "late": LateFusion, "hybrid": HybridFusion}
    if fusion_type not in fusion_classes:
        raise ValueError(f"Unknown fusion type: {fusion_type}")                          # :696
    return fusion_classes[fusion_type](modality_dims=modality_dims, num_classes=num_classes, **kwargs)
