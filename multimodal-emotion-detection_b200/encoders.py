"""Drop-in modules for src/encoders.py: SequenceEncoder (lstm), FrameEncoder (attention pooling),
build_encoder.  Same constructor / forward signatures, parameter names, shapes and default init as the
reference (the torch.nn modules below are parameter containers: their forward is never called), so a
reference state-dict loads with strict=True.  forward runs the sm_100a kernels of libavhot.so; inference
only (no autograd graph is recorded).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn as nn

from . import _lib, _rng
from ._handles import Handle, Workspace

_ws = Workspace()
_ws_split = Workspace()
_ws_frames = Workspace()


class SequenceEncoder(nn.Module):
    """src/encoders.py:15-192.  encoder_type 'lstm' is the hot path (configs/base.yaml:32-38); 'gru', 'cnn' and
    'transformer'  run on the fp32 kernels of csrc/gru.cu / csrc/cnn.cu / csrc/seqattn.cu; with
    precision = PREC_TC the transformer's nn.Linear layers run on the tcgen05 TF32 GEMM."""

    def __init__(self, input_dim: int, hidden_dim: int = 256, output_dim: int = 128, num_layers: int = 2,
                 encoder_type: str = 'lstm', dropout: float = 0.1):
        super().__init__()
        self.encoder_type = encoder_type
        self.hidden_dim = hidden_dim
        self.output_dim = output_dim
        self.input_dim = input_dim
        self.num_layers = num_layers
        self.precision = _lib.PREC_FP32
        if encoder_type == 'lstm':
            self.rnn = nn.LSTM(input_size=input_dim, hidden_size=hidden_dim, num_layers=num_layers,
                               batch_first=True, dropout=dropout if num_layers > 1 else 0.0,
                               bidirectional=False)                                  # src/encoders.py:54-61
            self.projection = nn.Linear(hidden_dim, output_dim)                      # :62
        elif encoder_type == 'gru':
            self.rnn = nn.GRU(input_size=input_dim, hidden_size=hidden_dim, num_layers=num_layers,
                              batch_first=True, dropout=dropout if num_layers > 1 else 0.0,
                              bidirectional=False)                                   # src/encoders.py:68-75
            self.projection = nn.Linear(hidden_dim, output_dim)                      # :76
        elif encoder_type == 'cnn':
            self.conv1 = nn.Conv1d(input_dim, hidden_dim, kernel_size=5, padding=2)  # src/encoders.py:86-93
            self.bn1 = nn.BatchNorm1d(hidden_dim)
            self.conv2 = nn.Conv1d(hidden_dim, hidden_dim, kernel_size=3, padding=1)
            self.bn2 = nn.BatchNorm1d(hidden_dim)
            self.act = nn.ReLU(inplace=True)
            self.dropout_layer = nn.Dropout(p=dropout)
            self.pool = nn.AdaptiveAvgPool1d(1)
            self.projection = nn.Linear(hidden_dim, output_dim)
        elif encoder_type == 'transformer':
            self.input_proj = nn.Linear(input_dim, hidden_dim)                       # src/encoders.py:100-112
            encoder_layer = nn.TransformerEncoderLayer(d_model=hidden_dim, nhead=4, dim_feedforward=hidden_dim * 4,
                                                       dropout=dropout, batch_first=False, activation='gelu')
            self.transformer = nn.TransformerEncoder(encoder_layer, num_layers=num_layers, enable_nested_tensor=False)
            self.max_len = 4096
            self.pos_embedding = nn.Embedding(self.max_len, hidden_dim)
            self.projection = nn.Linear(hidden_dim, output_dim)
        else:
            raise ValueError(f"Unknown encoder type: {encoder_type}")                # :115
        self._handle = Handle({"gru": "avhot_gru_destroy", "cnn": "avhot_cnn_destroy",
                               "transformer": "avhot_tfm_destroy"}.get(encoder_type, "avhot_lstm_destroy"))

    # ---- transformer (csrc/seqattn.cu) -----------------------------------------------------------------------------
    def _tfm_tensors(self):
        ts = [self.input_proj.weight, self.input_proj.bias, self.pos_embedding.weight, self.projection.weight,
              self.projection.bias]
        for layer in self.transformer.layers:
            ts += [layer.self_attn.in_proj_weight, layer.self_attn.in_proj_bias, layer.self_attn.out_proj.weight,
                   layer.self_attn.out_proj.bias, layer.linear1.weight, layer.linear1.bias, layer.linear2.weight,
                   layer.linear2.bias, layer.norm1.weight, layer.norm1.bias, layer.norm2.weight, layer.norm2.bias]
        return ts

    def _get_tfm_handle(self, dev):
        src = self._tfm_tensors()

        def create():
            ts = [_lib.f32c(t.detach()) for t in src]
            _lib.require_cuda(*ts)
            layer0 = self.transformer.layers[0]
            if len(self.transformer.layers) > _lib.TFM_MAX_LAYERS:
                raise ValueError(f"SequenceEncoder(transformer): at most {_lib.TFM_MAX_LAYERS} layers are built")
            if layer0.norm_first or self.transformer.norm is not None:
                raise ValueError("SequenceEncoder(transformer): only the post-norm layer the reference builds")
            d = _lib.TfmDesc()
            d.input_dim, d.hidden_dim, d.output_dim = self.input_dim, self.hidden_dim, self.output_dim
            d.num_layers, d.num_heads = len(self.transformer.layers), layer0.self_attn.num_heads
            d.ffn_dim, d.max_len, d.ln_eps = layer0.linear1.out_features, self.max_len, float(layer0.norm1.eps)
            d.input_w, d.input_b, d.pos, d.proj_w, d.proj_b = [t.data_ptr() for t in ts[:5]]
            names = ("in_proj_w", "in_proj_b", "out_w", "out_b", "lin1_w", "lin1_b", "lin2_w", "lin2_b", "n1_w", "n1_b",
                     "n2_w", "n2_b")
            for l in range(d.num_layers):
                for j, n in enumerate(names):
                    getattr(d, n)[l] = ts[5 + 12 * l + j].data_ptr()
            out = C.c_void_p()
            _lib.check(_lib.load().avhot_tfm_create(C.byref(d), C.byref(out), _lib.stream_ptr(dev)),
                       "SequenceEncoder(transformer)")
            torch.cuda.current_stream(dev).synchronize()       # the handle copies from `ts` on the stream
            return out.value

        return self._handle.get(src, create)

    def _tfm_drop_p(self) -> float:
        # the six nn.Dropout modules of the layers share p; MCDropoutUncertainty flips all of them (src/uncertainty.py:36-46)
        mods = [m for layer in self.transformer.layers for m in (layer.dropout1, layer.dropout, layer.dropout2)]
        if len({m.training for m in mods}) != 1:
            raise ValueError("SequenceEncoder(transformer): the layers' dropout modules must be in the same mode")
        return float(mods[0].p) if mods[0].training else 0.0

    @torch.no_grad()
    def _forward_transformer(self, x: torch.Tensor, lengths=None, passes: int = 1, p: Optional[float] = None,
                             seed: Optional[int] = None, clip0: Optional[int] = None,
                             t0: Optional[int] = None) -> torch.Tensor:
        """(B, T, in) -> (B, passes, out).  src/encoders.py:170-192; pass t draws the masks of (seed, t0 + t)."""
        dev = x.device
        B, T, _ = x.shape
        if p is None:
            p = self._tfm_drop_p()
        cs, ct, cc = _rng.current()
        seed = cs if seed is None else seed
        t0 = ct if t0 is None else t0
        clip0 = cc if clip0 is None else clip0
        z = torch.empty(B, passes, self.output_dim, device=dev, dtype=torch.float32)
        if B == 0:
            return z
        lib = _lib.load()
        with torch.cuda.device(dev):
            h = self._get_tfm_handle(dev)
            ws = _ws.get(lib.avhot_tfm_workspace_bytes(h, B, T, passes), dev)
            prec = _lib.PREC_TC if self.precision == _lib.PREC_TC else _lib.PREC_FP32
            _lib.check(lib.avhot_tfm_forward(h, _lib.ptr(x), _lib.ptr(lengths), _lib.ptr(z), B, T, passes, float(p), seed,
                                             clip0, t0, prec, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)),
                       "SequenceEncoder.forward (transformer)")
        return z

    def _cnn_tensors(self):
        return [self.conv1.weight, self.conv1.bias, self.bn1.weight, self.bn1.bias, self.bn1.running_mean,
                self.bn1.running_var, self.conv2.weight, self.conv2.bias, self.bn2.weight, self.bn2.bias,
                self.bn2.running_mean, self.bn2.running_var, self.projection.weight, self.projection.bias]

    def _get_cnn_handle(self, dev):
        src = self._cnn_tensors()

        def create():
            ts = [_lib.f32c(t.detach()) for t in src]
            _lib.require_cuda(*ts)
            d = _lib.CnnDesc()
            d.input_dim, d.hidden_dim, d.output_dim = self.input_dim, self.hidden_dim, self.output_dim
            (d.conv1_weight, d.conv1_bias, d.bn1_weight, d.bn1_bias, d.bn1_mean, d.bn1_var, d.conv2_weight,
             d.conv2_bias, d.bn2_weight, d.bn2_bias, d.bn2_mean, d.bn2_var, d.proj_weight, d.proj_bias) = \
                [t.data_ptr() for t in ts]
            d.bn_eps = float(self.bn1.eps)
            out = C.c_void_p()
            _lib.check(_lib.load().avhot_cnn_create(C.byref(d), C.byref(out), _lib.stream_ptr(dev)),
                       "SequenceEncoder(cnn)")
            torch.cuda.current_stream(dev).synchronize()       # the handle folds / copies from `ts` on the stream
            return out.value

        return self._handle.get(src, create)

    @torch.no_grad()
    def _forward_cnn(self, x: torch.Tensor, passes: int = 1, p: Optional[float] = None, seed: Optional[int] = None,
                     clip0: Optional[int] = None, t0: Optional[int] = None) -> torch.Tensor:
        """(B, T, in) -> (B, passes, out).  src/encoders.py:157-168; BatchNorm uses its running statistics (the
        inference path: MCDropoutUncertainty puts only nn.Dropout modules in train mode, src/uncertainty.py:36-46)."""
        if self.bn1.training or self.bn2.training:
            raise ValueError("SequenceEncoder(cnn): training-mode BatchNorm (batch statistics) is not built; call eval()")
        if float(self.bn1.eps) != float(self.bn2.eps):
            raise ValueError("SequenceEncoder(cnn): bn1 and bn2 must share eps")
        dev = x.device
        B, T, _ = x.shape
        if p is None:
            p = float(self.dropout_layer.p) if self.dropout_layer.training else 0.0
        cs, ct, cc = _rng.current()
        seed = cs if seed is None else seed
        t0 = ct if t0 is None else t0
        clip0 = cc if clip0 is None else clip0
        z = torch.empty(B, passes, self.output_dim, device=dev, dtype=torch.float32)
        if B == 0:
            return z
        lib = _lib.load()
        with torch.cuda.device(dev):
            h = self._get_cnn_handle(dev)
            ws = _ws.get(lib.avhot_cnn_workspace_bytes(h, B, T, passes), dev)
            _lib.check(lib.avhot_cnn_forward(h, _lib.ptr(x), _lib.ptr(z), B, T, passes, float(p), seed, clip0, t0,
                                             _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)),
                       "SequenceEncoder.forward (cnn)")
        return z

    def _params(self):
        ps = []
        for l in range(self.num_layers):
            ps += [getattr(self.rnn, f"weight_ih_l{l}"), getattr(self.rnn, f"weight_hh_l{l}"),
                   getattr(self.rnn, f"bias_ih_l{l}"), getattr(self.rnn, f"bias_hh_l{l}")]
        return ps + [self.projection.weight, self.projection.bias]

    def _get_handle(self, dev):
        ps = [_lib.f32c(p.detach()) for p in self._params()]
        _lib.require_cuda(*ps)

        def create():
            d = _lib.LstmDesc()
            d.input_dim, d.hidden_dim = self.input_dim, self.hidden_dim
            d.num_layers, d.output_dim = self.num_layers, self.output_dim
            for l in range(self.num_layers):
                d.weight_ih[l], d.weight_hh[l] = ps[4 * l].data_ptr(), ps[4 * l + 1].data_ptr()
                d.bias_ih[l], d.bias_hh[l] = ps[4 * l + 2].data_ptr(), ps[4 * l + 3].data_ptr()
            d.proj_weight, d.proj_bias = ps[-2].data_ptr(), ps[-1].data_ptr()
            out = C.c_void_p()
            create_fn = _lib.load().avhot_gru_create if self.encoder_type == 'gru' else _lib.load().avhot_lstm_create
            _lib.check(create_fn(C.byref(d), C.byref(out), _lib.stream_ptr(dev)), "SequenceEncoder")
            if self.encoder_type == 'gru':
                torch.cuda.current_stream(dev).synchronize()   # the handle copies from temporaries on the stream
            return out.value

        return self._handle.get(self._params(), create)

    @torch.no_grad()
    def forward(self, sequence: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
        dev = _lib.require_cuda(sequence)
        _lib.check_device(dev)
        if sequence.dim() != 3 or sequence.shape[2] != self.input_dim:
            raise ValueError(f"sequence must be (B, T, {self.input_dim}), got {tuple(sequence.shape)}")
        x = _lib.f32c(sequence)
        B, T, _ = x.shape
        if T < 1:
            raise ValueError("empty sequence")
        if lengths is not None:
            # values outside [1, T] are clamped by the kernels (the reference's pack_padded_sequence raises for them;
            # checking here would cost a host round trip per call)
            lengths = torch.as_tensor(lengths).to(device=dev, dtype=torch.int32).contiguous()
            if lengths.shape != (B,):
                raise ValueError("lengths must be (B,)")
        if self.encoder_type == 'cnn':
            return self._forward_cnn(x)[:, 0, :]               # `lengths` is ignored for this type (:157-168)
        if self.encoder_type == 'transformer':
            return self._forward_transformer(x, lengths)[:, 0, :]
        if self.rnn.training and self.rnn.dropout > 0 and self.num_layers > 1:
            # nn.LSTM / nn.GRU apply inter-layer dropout in train mode; these kernels are the inference path (under
            # MCDropoutUncertainty the rnn stays in eval mode: it is not an nn.Dropout module, src/uncertainty.py:41-46)
            raise ValueError(f"SequenceEncoder({self.encoder_type}): train-mode inter-layer dropout is not built; "
                             "call eval() (MC-dropout passes keep the recurrent layers in eval mode)")
        out = torch.empty(B, self.output_dim, device=dev, dtype=torch.float32)
        if B == 0:
            return out
        lib = _lib.load()
        with torch.cuda.device(dev):
            h = self._get_handle(dev)
            if self.encoder_type == 'gru':
                ws = _ws.get(lib.avhot_gru_workspace_bytes(h, B, T), dev)
                _lib.check(lib.avhot_gru_forward(h, _lib.ptr(x), _lib.ptr(lengths), _lib.ptr(out), B, T, _lib.ptr(ws),
                                                 ws.numel(), _lib.stream_ptr(dev)), "SequenceEncoder.forward (gru)")
                return out
            nbytes = lib.avhot_lstm_workspace_bytes(h, B, T, self.precision)
            ws = _ws.get(nbytes, dev)
            _lib.check(lib.avhot_lstm_forward(h, _lib.ptr(x), _lib.ptr(lengths), _lib.ptr(out), B, T,
                                              self.precision, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)),
                       "SequenceEncoder.forward")
        return out


    @torch.no_grad()
    def forward_from_waveform(self, wave: torch.Tensor) -> torch.Tensor:
        """Tensor-core pipeline, front-end fused with the LSTM's operand preparation: (B, N) waveform -> log-mel/MFCC
        front-end writing the fp16 split operand of the first input GEMM directly (avhot_frontend_ex) -> LSTM
        (avhot_lstm_forward_split).  Same result as forward(logmel_mfcc(wave)) with precision 'tc'."""
        from . import dataprocessing as dp
        dev = _lib.require_cuda(wave)
        _lib.check_device(dev)
        if wave.dim() == 3 and wave.shape[-1] == 1:
            wave = wave[..., 0]
        if wave.dim() != 2 or self.input_dim != dp.N_FEATS or self.encoder_type != 'lstm':
            raise ValueError("forward_from_waveform needs a (B, N) waveform, input_dim == 104 and encoder_type 'lstm'")
        wave = _lib.f32c(wave)
        B, N = wave.shape
        T = dp.n_frames(N)
        out = torch.empty(B, self.output_dim, device=dev, dtype=torch.float32)
        if B == 0:
            return out
        lib = _lib.load()
        with torch.cuda.device(dev):
            h = self._get_handle(dev)
            Bp = (B + 127) // 128 * 128
            pitch = lib.avhot_lstm_split_pitch(h)
            xs = _ws_split.get(2 * T * Bp * pitch, dev)
            _lib.check(lib.avhot_frontend_ex(_lib.ptr(wave), None, _lib.ptr(xs), B, Bp, pitch, N, _lib.stream_ptr(dev)),
                       "logmel_mfcc (fused)")
            ws = _ws.get(lib.avhot_lstm_workspace_bytes(h, B, T, _lib.PREC_TC), dev)
            # the operand is the front-end's: its MFCC columns are the DCT of its log-mel columns, which layer 0 exploits
            _lib.check(lib.avhot_lstm_forward_frontend(h, _lib.ptr(xs), None, _lib.ptr(out), B, T, _lib.ptr(ws), ws.numel(),
                                                       _lib.stream_ptr(dev)), "SequenceEncoder.forward (fused)")
        return out


class FrameEncoder(nn.Module):
    """src/encoders.py:196-327.  temporal_pooling 'attention' is the hot path (configs/base.yaml:43)."""

    def __init__(self, frame_dim: int, hidden_dim: int = 256, output_dim: int = 128,
                 temporal_pooling: str = 'attention', dropout: float = 0.1):
        super().__init__()
        self.temporal_pooling = temporal_pooling
        self.frame_dim, self.hidden_dim, self.output_dim = frame_dim, hidden_dim, output_dim
        self.precision = _lib.PREC_FP32
        self.frame_mlp = nn.Sequential(nn.Linear(frame_dim, hidden_dim), nn.ReLU(inplace=True),
                                       nn.Dropout(dropout))                          # :225-229
        self._dropout = nn.Dropout(dropout)                                           # :230
        if temporal_pooling == 'attention':
            self.attention = nn.Linear(hidden_dim, 1)                                 # :236
        elif temporal_pooling in ('average', 'max'):
            pass                                                                      # :237-240: no extra parameters
        else:
            raise ValueError(f"Unknown pooling: {temporal_pooling}")                  # :242
        self.projection = nn.Sequential(nn.LayerNorm(hidden_dim), nn.Linear(hidden_dim, output_dim))  # :246-249
        self._handle = Handle("avhot_frame_destroy")

    def _params(self):
        att = [self.attention.weight, self.attention.bias] if self.temporal_pooling == 'attention' else []
        return [self.frame_mlp[0].weight, self.frame_mlp[0].bias] + att + [
            self.projection[0].weight, self.projection[0].bias, self.projection[1].weight, self.projection[1].bias]

    def _get_handle(self, dev):
        ps = [_lib.f32c(p.detach()) for p in self._params()]
        _lib.require_cuda(*ps)

        def create():
            d = _lib.FrameDesc()
            d.frame_dim, d.hidden_dim, d.output_dim = self.frame_dim, self.hidden_dim, self.output_dim
            ptrs = [p.data_ptr() for p in ps]
            if self.temporal_pooling != 'attention':
                ptrs[2:2] = [None, None]                      # no attention layer (average / max pooling)
            (d.mlp_weight, d.mlp_bias, d.attn_weight, d.attn_bias, d.ln_weight, d.ln_bias, d.proj_weight,
             d.proj_bias) = ptrs
            out = C.c_void_p()
            _lib.check(_lib.load().avhot_frame_create(C.byref(d), C.byref(out), _lib.stream_ptr(dev)),
                       "FrameEncoder")
            return out.value

        return self._handle.get(self._params(), create)

    # -- the two halves of forward, exposed so the MC wrapper can hoist the pass-invariant one ------------
    @torch.no_grad()
    def hidden(self, frames: torch.Tensor) -> torch.Tensor:
        """frame_mlp[0:2] = Linear + ReLU (:225-227): (B, F, frame_dim) fp32, or uint8 pixels k meaning k/255
        (src/dataprocessing.py:261-265) -> (B, F, H), before dropout."""
        dev = _lib.require_cuda(frames)
        _lib.check_device(dev)
        if frames.dim() != 3 or frames.shape[2] != self.frame_dim:
            raise ValueError(f"frames must be (B, F, {self.frame_dim}), got {tuple(frames.shape)}")
        u8 = frames.dtype == torch.uint8
        x = frames.contiguous() if u8 else _lib.f32c(frames)
        B, F, _ = x.shape
        if F < 1:
            raise ValueError("no frames")
        hid = torch.empty(B, F, self.hidden_dim, device=dev, dtype=torch.float32)
        if B == 0:
            return hid
        lib = _lib.load()
        with torch.cuda.device(dev):
            h = self._get_handle(dev)
            nbytes = lib.avhot_frame_workspace_bytes(h, B, F, self.precision) if u8 else 0
            ws = _ws_frames.get(nbytes, dev)
            _lib.check(lib.avhot_frame_hidden(h, None if u8 else _lib.ptr(x), _lib.ptr(x) if u8 else None,
                                              _lib.ptr(hid), B, F, self.precision, _lib.ptr(ws), ws.numel(),
                                              _lib.stream_ptr(dev)), "FrameEncoder.hidden")
        return hid

    @torch.no_grad()
    def tail(self, hid: torch.Tensor, mask: Optional[torch.Tensor] = None, passes: int = 1, p: float = 0.0,
             seed: int = 0, clip0: int = 0, t0: int = 0, project: bool = True, clip_ids=None) -> torch.Tensor:
        """Everything after the ReLU for `passes` stochastic passes: (B, F, H) -> (B, passes, out).
        project=False stops after projection[0] (the LayerNorm) and returns (B, passes, H): the tensor-core hot path
        folds projection[1] into HybridFusion's proj_layers.video (HybridFusion.run(..., video_norm_from=...))."""
        dev = _lib.require_cuda(hid, mask)
        B, F, _ = hid.shape
        valid = None
        if mask is not None:
            valid = torch.as_tensor(mask).to(device=dev).bool().to(torch.uint8).contiguous()
            if valid.shape != (B, F):
                raise ValueError("mask must be (B, num_frames)")
        z = torch.empty(B, passes, self.output_dim if project else self.hidden_dim, device=dev, dtype=torch.float32)
        if B == 0:
            return z
        with torch.cuda.device(dev):
            h = self._get_handle(dev)
            if clip_ids is not None:                      # scattered clips: per-row global indices (fp32 path only)
                if not project:
                    raise ValueError("clip_ids needs project=True")
                _lib.check(_lib.load().avhot_frame_tail_ids(h, _lib.ptr(hid), _lib.ptr(valid), _lib.ptr(z), B, F, passes,
                                                            float(p), seed, _lib.ptr(clip_ids), t0, _lib.stream_ptr(dev)),
                           "FrameEncoder.tail")
                return z
            if self.temporal_pooling != 'attention':
                if not project:
                    raise ValueError("project=False is the attention-pooling hot path")
                pooling = _lib.POOL_AVERAGE if self.temporal_pooling == 'average' else _lib.POOL_MAX
                _lib.check(_lib.load().avhot_frame_tail_pool(h, _lib.ptr(hid), _lib.ptr(valid), _lib.ptr(z), B, F, passes,
                                                             float(p), seed, clip0, t0, pooling, _lib.stream_ptr(dev)),
                           "FrameEncoder.tail")
                return z
            fn = _lib.load().avhot_frame_tail if project else _lib.load().avhot_frame_tail_norm
            _lib.check(fn(h, _lib.ptr(hid), _lib.ptr(valid), _lib.ptr(z), B, F, passes, float(p), seed, clip0, t0,
                          _lib.stream_ptr(dev)), "FrameEncoder.tail")
        return z

    def _drop_p(self) -> float:
        # nn.Dropout is active only in train mode; MCDropoutUncertainty flips exactly these two modules
        # (src/uncertainty.py:36-46).  Both sites share p (one ctor argument, src/encoders.py:228,230).
        d0, d1 = self.frame_mlp[2], self._dropout
        if d0.training != d1.training:
            raise ValueError("frame_mlp dropout and _dropout must be in the same mode")
        return float(d0.p) if d0.training else 0.0

    def forward(self, frames: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        hid = self.hidden(frames)
        seed, t, clip0 = _rng.current()
        return self.tail(hid, mask, 1, self._drop_p(), seed, clip0, t)[:, 0, :]

    @torch.no_grad()
    def attention_pool(self, frames: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """src/encoders.py:300-327 on already-encoded frames (B, F, H) -> (B, H): the pooling stage of the tail kernel
        on its own (avhot_frame_attention_pool)."""
        dev = _lib.require_cuda(frames, mask)
        _lib.check_device(dev)
        if self.temporal_pooling != 'attention':
            raise AttributeError("attention_pool needs temporal_pooling='attention'")    # the reference has no self.attention
        if frames.dim() != 3 or frames.shape[2] != self.hidden_dim:
            raise ValueError(f"frames must be (B, F, {self.hidden_dim}), got {tuple(frames.shape)}")
        x = _lib.f32c(frames)
        B, F, _ = x.shape
        valid = None
        if mask is not None:
            valid = torch.as_tensor(mask).to(device=dev).bool().to(torch.uint8).contiguous()
            if valid.shape != (B, F):
                raise ValueError("mask must be (B, num_frames)")
        pooled = torch.empty(B, self.hidden_dim, device=dev, dtype=torch.float32)
        if B > 0:
            with torch.cuda.device(dev):
                h = self._get_handle(dev)
                _lib.check(_lib.load().avhot_frame_attention_pool(h, _lib.ptr(x), _lib.ptr(valid), _lib.ptr(pooled), B, F,
                                                                  _lib.stream_ptr(dev)), "FrameEncoder.attention_pool")
        return pooled


def build_encoder(modality: str, input_dim: int, output_dim: int, encoder_config: dict = None) -> nn.Module:
    """src/encoders.py:532-636: same selection rules; 'mlp' and 'pretrained_cnn' are out of scope."""
    if encoder_config is None:
        encoder_config = {}
    cfg = dict(encoder_config)
    enc_type = cfg.pop("type", None)
    in_dim = cfg.pop("input_dim", input_dim)
    if enc_type is None:
        mod_lower = modality.lower()
        if mod_lower in {"video", "frames"}:
            enc_type = "frame"
        elif mod_lower in {"imu", "mocap", "audio", "accelerometer", "gyro", "magnetometer", "imu_hand",
                           "imu_chest", "imu_ankle"}:
            enc_type = "sequence"
        else:
            enc_type = "mlp"
    if enc_type == "frame":
        temporal_pooling = cfg.pop("temporal_pooling", "attention")
        hidden_dim = cfg.pop("hidden_dim", None)
        dropout = cfg.pop("dropout", 0.1)
        return FrameEncoder(frame_dim=in_dim, hidden_dim=hidden_dim if hidden_dim is not None else output_dim * 2,
                            output_dim=output_dim, temporal_pooling=temporal_pooling, dropout=dropout)
    if enc_type == "sequence":
        encoder_kind = cfg.pop("encoder_type", "lstm")
        num_layers = cfg.pop("num_layers", 2)
        hidden_dim = cfg.pop("hidden_dim", None)
        dropout = cfg.pop("dropout", 0.1)
        return SequenceEncoder(input_dim=in_dim, hidden_dim=hidden_dim if hidden_dim is not None else output_dim * 2,
                               output_dim=output_dim, num_layers=num_layers, encoder_type=encoder_kind,
                               dropout=dropout)
    if enc_type in ("mlp", "pretrained_cnn"):
        raise ValueError(f"encoder type '{enc_type}' exists in the reference but is not built in this library")
    raise ValueError(f"Unknown encoder type '{enc_type}' for modality '{modality}'")
