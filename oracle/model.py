"""Encoders -> cross-modal attention -> fusion -> MC-dropout, restated on CPU.
TEST INFRASTRUCTURE (see oracle/__init__.py).

Plain torch tensor arithmetic on CPU (matmul / exp / tanh / mean), written as
functions over a flat state-dict (SURVEY.md Appendix A key names), so the same
code runs in fp32 (like the reference) or fp64 (tight checker for the CUDA
kernels).  No nn.Module, no nn.LSTM: the recurrence is written out.  Each
function cites the reference lines it follows.  Validated against the
reference's own modules by tests/golden/make_golden.py (run in the build
container, where /root/reference exists); its outputs are the committed golden
vectors that tests/test_oracle_model.py re-checks everywhere.

Dropout: the reference draws masks from torch's global RNG; here every
``nn.Dropout`` site takes its mask from the Philox spec (oracle/philox.py)
through a ``DropCtx``.  ``ctx=None`` means eval mode (dropout is identity).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import philox as px


# bench.py's CPU arm sets this to a dict so that the Philox masks (numpy restatement of the spec, far slower than the
# native RNG the reference's nn.Dropout draws from) are generated once, outside its timed region, and looked up inside
MASK_CACHE: Optional[dict] = None


@dataclass
class DropCtx:
    """Which masks a stochastic pass uses: (seed, pass index t, global clip ids)."""
    seed: int
    t: int
    clip_ids: np.ndarray           # (B,) global clip indices
    p_encoder: float = 0.1         # FrameEncoder dropout (configs/base.yaml:46)
    p_fusion: float = 0.3          # model.dropout (configs/base.yaml:27)

    def apply(self, x: torch.Tensor, site: int, p: float) -> torch.Tensor:
        """nn.Dropout forward in train mode with a spec mask: x * keep / (1 - p)."""
        if p <= 0.0:
            return x
        B = x.shape[0]
        n = int(np.prod(x.shape[1:]))
        key = (self.seed, self.t, site, float(p), n, self.clip_ids.tobytes()) if MASK_CACHE is not None else None
        keep = MASK_CACHE.get(key) if key is not None else None
        if keep is None:
            keep = torch.from_numpy(px.dropout_keep(n, self.clip_ids, self.t, site, p, self.seed).reshape(x.shape))
            if key is not None:
                MASK_CACHE[key] = keep
        scale = torch.tensor(float(px.keep_scale(p)), dtype=torch.float32).to(x.dtype)
        return x * (keep.to(x.dtype) * scale)


def _lin(x, sd, name):
    return x @ sd[name + ".weight"].T + sd[name + ".bias"]


def _layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)      # biased, like nn.LayerNorm
    return (x - mu) / torch.sqrt(var + eps) * w + b


def _softmax(x, dim=-1):
    m = x.max(dim=dim, keepdim=True).values
    e = torch.exp(x - m)
    return e / e.sum(dim=dim, keepdim=True)


def sequence_encoder_lstm(seq, sd, prefix="encoders.audio.", num_layers=2):
    """SequenceEncoder(encoder_type='lstm').forward, lengths=None branch.

    src/encoders.py:54-62 (nn.LSTM batch_first, unidirectional, zero initial
    state), :148-155 (h_n[-1] -> projection).  nn.LSTM cell: gate order
    i, f, g, o in the 4H rows; c' = f*c + i*g; h' = o*tanh(c').  Inter-layer
    dropout is inactive in eval and under MC-dropout.
    """
    B, T, _ = seq.shape
    x = seq
    h = None
    for l in range(num_layers):
        w_ih = sd[f"{prefix}rnn.weight_ih_l{l}"]
        w_hh = sd[f"{prefix}rnn.weight_hh_l{l}"]
        bias = sd[f"{prefix}rnn.bias_ih_l{l}"] + sd[f"{prefix}rnn.bias_hh_l{l}"]
        H = w_hh.shape[1]
        xp = x @ w_ih.T + bias                      # (B, T, 4H): input part, all steps
        h = torch.zeros(B, H, dtype=seq.dtype)
        c = torch.zeros(B, H, dtype=seq.dtype)
        outs = []
        for t in range(T):
            g = xp[:, t] + h @ w_hh.T
            i_g = torch.sigmoid(g[:, 0 * H:1 * H])
            f_g = torch.sigmoid(g[:, 1 * H:2 * H])
            g_g = torch.tanh(g[:, 2 * H:3 * H])
            o_g = torch.sigmoid(g[:, 3 * H:4 * H])
            c = f_g * c + i_g * g_g
            h = o_g * torch.tanh(c)
            outs.append(h)
        x = torch.stack(outs, dim=1)
    return _lin(h, sd, prefix + "projection")


def sequence_encoder_gru(seq, sd, prefix="gru_enc.", num_layers=2, lengths=None):
    """SequenceEncoder(encoder_type='gru').forward.  src/encoders.py:68-76 (nn.GRU batch_first, unidirectional, zero
    initial state), :137-155 (h_n[-1] -> projection; with `lengths` h_n is the state after each clip's last valid
    step, the packed-sequence branch :139-147).  nn.GRU cell, gate rows r, z, n:
    r = sig(W_ir x + b_ir + W_hr h + b_hr), z likewise, n = tanh(W_in x + b_in + r * (W_hn h + b_hn)),
    h' = (1 - z) * n + z * h."""
    B, T, _ = seq.shape
    x = seq
    last = None
    for l in range(num_layers):
        w_ih, w_hh = sd[f"{prefix}rnn.weight_ih_l{l}"], sd[f"{prefix}rnn.weight_hh_l{l}"]
        b_ih, b_hh = sd[f"{prefix}rnn.bias_ih_l{l}"], sd[f"{prefix}rnn.bias_hh_l{l}"]
        H = w_hh.shape[1]
        xp = x @ w_ih.T + b_ih
        h = torch.zeros(B, H, dtype=seq.dtype)
        last = torch.zeros(B, H, dtype=seq.dtype)
        outs = []
        for t in range(T):
            hp = h @ w_hh.T + b_hh
            r = torch.sigmoid(xp[:, t, 0 * H:1 * H] + hp[:, 0 * H:1 * H])
            z = torch.sigmoid(xp[:, t, 1 * H:2 * H] + hp[:, 1 * H:2 * H])
            n = torch.tanh(xp[:, t, 2 * H:3 * H] + r * hp[:, 2 * H:3 * H])
            h = (1 - z) * n + z * h
            outs.append(h)
            if lengths is not None:
                sel = (torch.as_tensor(lengths).clamp(1, T) - 1) == t
                last[sel] = h[sel]
        if lengths is None:
            last = h
        x = torch.stack(outs, dim=1)
    return _lin(last, sd, prefix + "projection")


def sequence_encoder_cnn(seq, sd, prefix="cnn_enc.", ctx: Optional[DropCtx] = None, eps=1e-5):
    """SequenceEncoder(encoder_type='cnn').forward.  src/encoders.py:157-168: transpose to (B, C, T), Conv1d(k 5, pad 2),
    BatchNorm1d (eval: running statistics), ReLU, Dropout, Conv1d(k 3, pad 1), BatchNorm1d, ReLU, mean over T, the same
    Dropout module again, Linear.  Dropout sites: CNN0 on the (B, H, T) activation (element h * T + t), CNN1 on the
    pooled (B, H) vector; p = the encoder's dropout (ctx.p_encoder)."""
    def conv1d(x, w, b, pad):                      # x (B, C, T), w (O, C, K): written out as K shifted matmuls
        B, C, T = x.shape
        xp = torch.nn.functional.pad(x, (pad, pad))
        out = b.view(1, -1, 1).expand(B, w.shape[0], T).clone()
        for k in range(w.shape[2]):
            out = out + torch.einsum("oc,bct->bot", w[:, :, k], xp[:, :, k:k + T])
        return out

    def bn(x, name):
        g = sd[f"{prefix}{name}.weight"] / torch.sqrt(sd[f"{prefix}{name}.running_var"] + eps)
        return (x - sd[f"{prefix}{name}.running_mean"].view(1, -1, 1)) * g.view(1, -1, 1) + sd[f"{prefix}{name}.bias"].view(1, -1, 1)

    x = seq.transpose(1, 2)
    x = torch.relu(bn(conv1d(x, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"], 2), "bn1"))
    if ctx is not None:
        x = ctx.apply(x, px.SITE_CNN0, ctx.p_encoder)
    x = torch.relu(bn(conv1d(x, sd[prefix + "conv2.weight"], sd[prefix + "conv2.bias"], 1), "bn2"))
    x = x.mean(dim=2)
    if ctx is not None:
        x = ctx.apply(x, px.SITE_CNN1, ctx.p_encoder)
    return _lin(x, sd, prefix + "projection")



# ---- transformer encoder, TemporalAttention, PairwiseModalityAttention ---------------
def _mha_self(x, valid, w_in, b_in, w_out, b_out, num_heads):
    """nn.MultiheadAttention(x, x, x, key_padding_mask=~valid) in eval mode (torch/nn/functional.py
    multi_head_attention_forward): packed in_proj (3H, H) = [q | k | v], scores scaled by 1/sqrt(head dim), padded
    keys -inf, softmax, out_proj.  x (B, T, H); valid (B, T) bool or None."""
    B, T, H = x.shape
    dh = H // num_heads
    qkv = x @ w_in.T + b_in
    q, k, v = (t.view(B, T, num_heads, dh).transpose(1, 2) for t in qkv.split(H, dim=-1))
    scores = (q @ k.transpose(-2, -1)) * (1.0 / math.sqrt(dh))
    if valid is not None:
        scores = scores.masked_fill(~valid[:, None, None, :], float("-inf"))
    attn = _softmax(scores, dim=-1)
    ctx = (attn @ v).transpose(1, 2).reshape(B, T, H)
    return ctx @ w_out.T + b_out


def _gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))       # F.gelu, approximate='none'


def sequence_encoder_transformer(seq, sd, prefix="tfm_enc.", num_layers=2, num_heads=4, lengths=None,
                                 ctx: Optional[DropCtx] = None, p: float = 0.1, max_len=4096):
    """SequenceEncoder(encoder_type='transformer').forward.  src/encoders.py:170-192 over the modules built at
    :100-112: input_proj, + pos_embedding[min(t, max_len - 1)], num_layers post-norm nn.TransformerEncoderLayer
    (nhead 4, dim_feedforward 4 H, GELU; x = norm1(x + dropout1(SA(x))), x = norm2(x + dropout2(linear2(dropout(
    gelu(linear1(x)))))), torch/nn/modules/transformer.py), key padding from `lengths`, mean over the valid steps,
    projection.  Dropout sites (only the nn.Dropout modules are stochastic under MC): SITE_TFM + 3 l + {0, 1, 2} with
    per-clip masks (T, H), (T, 4H), (T, H) - the reference holds these tensors as (T, B, .); the clip axis is 1 there."""
    B, T, _ = seq.shape
    x = _lin(seq, sd, prefix + "input_proj")
    pos_idx = torch.arange(T).clamp_max(max_len - 1)
    x = x + sd[prefix + "pos_embedding.weight"][pos_idx].unsqueeze(0)
    valid = None
    if lengths is not None:
        valid = torch.arange(T).unsqueeze(0) < torch.as_tensor(lengths).view(-1, 1)
    for l in range(num_layers):
        lp = f"{prefix}transformer.layers.{l}."
        sa = _mha_self(x, valid, sd[lp + "self_attn.in_proj_weight"], sd[lp + "self_attn.in_proj_bias"],
                       sd[lp + "self_attn.out_proj.weight"], sd[lp + "self_attn.out_proj.bias"], num_heads)
        if ctx is not None:
            sa = ctx.apply(sa, px.SITE_TFM + 3 * l, p)
        x = _layer_norm(x + sa, sd[lp + "norm1.weight"], sd[lp + "norm1.bias"])
        ff = _gelu(_lin(x, sd, lp + "linear1"))
        if ctx is not None:
            ff = ctx.apply(ff, px.SITE_TFM + 3 * l + 1, p)
        ff = _lin(ff, sd, lp + "linear2")
        if ctx is not None:
            ff = ctx.apply(ff, px.SITE_TFM + 3 * l + 2, p)
        x = _layer_norm(x + ff, sd[lp + "norm2.weight"], sd[lp + "norm2.bias"])
    if valid is not None:
        m = valid.to(x.dtype).unsqueeze(-1)
        pooled = (x * m).sum(dim=1) / m.sum(dim=1).clamp_min(1.0)                          # :186-190
    else:
        pooled = x.mean(dim=1)                                                              # :192
    return _lin(pooled, sd, prefix + "projection")


def temporal_attention(seq, mask, sd, prefix="tattn.", num_heads=4):
    """TemporalAttention.forward.  src/attention.py:201-243: q / k / v projections of the same sequence, scaled
    scores, masked keys -inf (:219-224), softmax, context, masked queries zeroed (:232-236), out_proj.  Neither
    nn.Dropout of the module is ever applied (:227, :240 test attributes that do not exist).
    seq (B, S, F); mask (B, S) 1 = valid or None -> (attended (B, S, H), weights (B, heads, S, S))."""
    B, S, _ = seq.shape
    q, k, v = (_lin(seq, sd, prefix + n) for n in ("q_proj", "k_proj", "v_proj"))
    H = q.shape[-1]
    dh = H // num_heads
    q, k, v = (t.view(B, S, num_heads, dh).permute(0, 2, 1, 3) for t in (q, k, v))
    logits = (q @ k.transpose(-2, -1)) * (dh ** -0.5)
    if mask is not None:
        km = torch.as_tensor(mask).bool()
        logits = logits.masked_fill(~km[:, None, None, :], float("-inf"))
    w = _softmax(logits, dim=-1)
    ctxv = w @ v
    if mask is not None:
        ctxv = ctxv * km[:, None, :, None]
    ctxv = ctxv.permute(0, 2, 1, 3).reshape(B, S, H)
    return _lin(ctxv, sd, prefix + "out_proj"), w


def temporal_pool_sequence(seq, weights):
    """TemporalAttention.pool_sequence.  src/attention.py:264-274."""
    imp = weights.mean(dim=1).mean(dim=1)
    imp = imp / (imp.sum(dim=1, keepdim=True) + 1e-9)
    return torch.bmm(imp.unsqueeze(1), seq).squeeze(1)


def pairwise_modality_attention(feats: Dict[str, torch.Tensor], mask, sd, modality_dims: Dict[str, int],
                                prefix="pair.", num_heads=4, ctx: Optional[DropCtx] = None, p: float = 0.1):
    """PairwiseModalityAttention.forward (the second class of that name, src/attention.py:421-548, which replaces the
    first at import).  For every unordered pair (a, b): a attends to b with b's availability as the key mask if any
    clip has b (:497-507), and vice versa (:509-517); per modality: LayerNorm(self_proj(x) + msg_dropout(sum of
    messages)) (:527-536), zeroed where the modality is missing (:539-541).  Dropout sites: the CrossModalAttention
    blocks SITE_ATTN + their index in cross_attn (ModuleDict order), msg_dropout SITE_PAIR_MSG + modality index.
    -> (attended {name: (B, H)}, attention maps {"a_to_b": (B, heads, 1, 1)})."""
    names = list(modality_dims.keys())
    first = next(iter(feats.values()))
    B = first.shape[0]
    if mask is not None:
        mask = torch.as_tensor(mask)
        avail = {m: (mask[:, i] > 0) if mask.dtype != torch.bool else mask[:, i] for i, m in enumerate(names)}
    else:
        avail = {m: torch.ones(B, dtype=torch.bool) for m in names}
    order = []
    for i in range(len(names)):
        for j in range(i + 1, len(names)):
            order += [f"{names[i]}_to_{names[j]}", f"{names[j]}_to_{names[i]}"]
    msgs = {m: [] for m in names}
    maps = {}
    for i in range(len(names)):
        for j in range(i + 1, len(names)):
            a, b = names[i], names[j]
            for tgt, src in ((a, b), (b, a)):
                if avail[src].any():
                    key = f"{tgt}_to_{src}"
                    o, att = cross_modal_attention(feats[tgt], feats[src], feats[src], (~avail[src]).view(B, 1), sd,
                                                   f"{prefix}cross_attn.{key}.", num_heads, ctx=ctx,
                                                   site=px.SITE_ATTN + order.index(key), p=p if ctx is not None else 0.0)
                    msgs[tgt].append(o)
                    maps[key] = att
    out = {}
    H = sd[prefix + "out_ln.weight"].shape[0]
    for i, m in enumerate(names):
        msg = torch.stack(msgs[m], dim=0).sum(dim=0) if msgs[m] else torch.zeros(B, H, dtype=first.dtype)
        if ctx is not None:
            msg = ctx.apply(msg, px.SITE_PAIR_MSG + i, p)
        agg = _layer_norm(_lin(feats[m], sd, f"{prefix}self_proj.{m}") + msg, sd[prefix + "out_ln.weight"],
                          sd[prefix + "out_ln.bias"])
        if not avail[m].all():
            agg = agg * avail[m].to(agg.dtype).unsqueeze(1)
        out[m] = agg
    return out, maps

_LSTM_CACHE = {}


def sequence_encoder_lstm_torch(seq, sd, prefix="encoders.audio.", num_layers=2):
    """Same function as sequence_encoder_lstm, computed by torch.nn.LSTM on CPU - the op the reference itself
    calls (src/encoders.py:54-61, :148-154).  Used for the CPU-baseline timing so the baseline pays what the
    reference pays (ATen's fused LSTM cell), not the cost of a Python time loop; checked against the written-out
    recurrence in tests/test_oracle_model.py."""
    key = (id(sd), prefix, num_layers, seq.dtype)
    rnn = _LSTM_CACHE.get(key)
    if rnn is None:
        H = sd[f"{prefix}rnn.weight_hh_l0"].shape[1]
        rnn = torch.nn.LSTM(input_size=sd[f"{prefix}rnn.weight_ih_l0"].shape[1], hidden_size=H,
                            num_layers=num_layers, batch_first=True, bidirectional=False).to(seq.dtype)
        rnn.load_state_dict({k[len(prefix) + 4:]: v.to(seq.dtype) for k, v in sd.items()
                             if k.startswith(prefix + "rnn.")}, strict=True)
        rnn.eval()
        _LSTM_CACHE.clear()
        _LSTM_CACHE[key] = rnn
    with torch.inference_mode():
        _, (h_n, _) = rnn(seq)
    return _lin(h_n[-1], sd, prefix + "projection")


def frame_hidden(frames, sd, prefix="encoders.video."):
    """frame_mlp[0:2]: Linear(frame_dim -> H) + ReLU, before dropout.  src/encoders.py:225-227,272."""
    return torch.relu(_lin(frames, sd, prefix + "frame_mlp.0"))


def attention_pool(x, sd, prefix="encoders.video.", valid=None):
    """FrameEncoder.attention_pool, src/encoders.py:300-327 (valid: (B, F) bool or None)."""
    scores = (x @ sd[prefix + "attention.weight"].T + sd[prefix + "attention.bias"]).squeeze(-1)
    if valid is not None:
        scores = scores.masked_fill(~valid, torch.finfo(torch.float32).min)              # :321-324 (fp32's finfo.min)
    w = _softmax(scores, dim=1)
    return (w.unsqueeze(-1) * x).sum(dim=1)


def frame_encoder_tail(hid, sd, prefix="encoders.video.", ctx: Optional[DropCtx] = None, pooling="attention",
                       valid=None):
    """FrameEncoder after the ReLU: dropout -> temporal pooling -> dropout -> LN -> Linear.

    src/encoders.py:228 (Dropout inside frame_mlp), :320-326 (scores, softmax over
    frames, einsum 'bt,bth->bh'; mask=None on the hot path, src/train.py:136), :276-293 ('average': masked mean with
    the count clamped to >= 1; 'max': invalid frames filled with finfo.min), :296 (_dropout), :246-249,297
    (LayerNorm + Linear).
    """
    x = hid
    if ctx is not None:
        x = ctx.apply(x, px.SITE_FRAME_MLP, ctx.p_encoder)
    if pooling == "attention":
        pooled = attention_pool(x, sd, prefix, valid)
    elif pooling == "average":
        if valid is not None:
            mf = valid.to(x.dtype).unsqueeze(-1)
            pooled = (x * mf).sum(dim=1) / mf.sum(dim=1).clamp_min(1.0)
        else:
            pooled = x.mean(dim=1)
    elif pooling == "max":
        xm = x if valid is None else x.masked_fill(~valid.unsqueeze(-1), torch.finfo(torch.float32).min)
        pooled = xm.max(dim=1).values
    else:
        raise ValueError(f"Unknown pooling: {pooling}")
    if ctx is not None:
        pooled = ctx.apply(pooled, px.SITE_FRAME_POOL, ctx.p_encoder)
    y = _layer_norm(pooled, sd[prefix + "projection.0.weight"], sd[prefix + "projection.0.bias"])
    return _lin(y, sd, prefix + "projection.1")


def frame_encoder(frames, sd, prefix="encoders.video.", ctx: Optional[DropCtx] = None):
    """FrameEncoder.forward(frames, mask=None), temporal_pooling='attention'.  src/encoders.py:253-297."""
    return frame_encoder_tail(frame_hidden(frames, sd, prefix), sd, prefix, ctx)


def cross_modal_attention(query, key, value, invalid, sd, prefix, num_heads=4,
                          ctx: Optional[DropCtx] = None, site: int = px.SITE_ATTN,
                          p: float = 0.0):
    """CrossModalAttention.forward.  src/attention.py:80-141.

    query (B,Tq,Dq), key/value (B,Tk,Dk); invalid (B,Tk) bool (True = masked key)
    or None.  Returns (out (B,Tq,D) or (B,D) when Tq == 1, attn (B,heads,Tq,Tk)).
    """
    if query.dim() == 2:
        query = query.unsqueeze(1)
    if key.dim() == 2:
        key = key.unsqueeze(1)
    if value.dim() == 2:
        value = value.unsqueeze(1)
    B, Tq, _ = query.shape
    Tk = key.shape[1]
    q = _layer_norm(query, sd[prefix + "q_in_ln.weight"], sd[prefix + "q_in_ln.bias"])   # :97
    k = _layer_norm(key, sd[prefix + "k_in_ln.weight"], sd[prefix + "k_in_ln.bias"])     # :98
    v = _layer_norm(value, sd[prefix + "v_in_ln.weight"], sd[prefix + "v_in_ln.bias"])   # :99
    Q = _lin(q, sd, prefix + "q_proj")
    K = _lin(k, sd, prefix + "k_proj")
    V = _lin(v, sd, prefix + "v_proj")
    D = Q.shape[-1]
    dh = D // num_heads
    Q = Q.view(B, Tq, num_heads, dh).transpose(1, 2)
    K = K.view(B, Tk, num_heads, dh).transpose(1, 2)
    V = V.view(B, Tk, num_heads, dh).transpose(1, 2)
    scores = (Q @ K.transpose(-2, -1)) * (1.0 / math.sqrt(dh))                          # :112
    if invalid is not None:
        scores = scores.masked_fill(invalid[:, None, None, :], -1e4)                    # :115-118
    attn = _softmax(scores, dim=-1)                                                     # :121
    if invalid is not None:
        all_masked = invalid.all(dim=-1)                                                # :124-128
        if all_masked.any():
            attn = attn.clone()
            attn[all_masked] = 0.0
    if ctx is not None:
        attn = ctx.apply(attn, site, p)                                                 # :130
    context = (attn @ V).transpose(1, 2).contiguous().view(B, Tq, D)                    # :133-136
    out = _lin(context, sd, prefix + "out_proj")                                        # :137
    if Tq == 1:
        out = out.squeeze(1)                                                            # :140
    return out, attn


def hybrid_fusion(feats: Dict[str, torch.Tensor], mask, sd, prefix="fusion_model.",
                  modalities: Sequence[str] = ("audio", "video"), num_heads=4,
                  ctx: Optional[DropCtx] = None, return_info=False):
    """HybridFusion.forward.  src/fusion.py:455-530 (all modalities present)."""
    ref = feats[modalities[0]]
    B = ref.shape[0]
    M = len(modalities)
    if mask is None:
        mask_f = torch.ones(B, M, dtype=torch.float32)
    else:
        mask_f = torch.as_tensor(mask).float()
    invalid = mask_f <= 0                                                               # :474
    Z = torch.stack([_lin(feats[m], sd, f"{prefix}proj_layers.{m}") for m in modalities], dim=1)
    Z = _layer_norm(Z, sd[prefix + "pre_ln.weight"], sd[prefix + "pre_ln.bias"])        # :490-491
    p = ctx.p_fusion if ctx is not None else 0.0
    att, attn_w = [], []
    for i, m in enumerate(modalities):
        out, a = cross_modal_attention(Z[:, i:i + 1, :], Z, Z, invalid, sd,
                                       f"{prefix}attn_layers.{m}.", num_heads, ctx,
                                       px.SITE_ATTN + i, p)                             # :497-500
        out = _layer_norm(out + Z[:, i, :], sd[prefix + "post_ln.weight"],
                          sd[prefix + "post_ln.bias"])                                  # :502
        att.append(out)
        attn_w.append(a)
    H_att = torch.stack(att, dim=1)                                                     # :508
    g = torch.relu(_lin(H_att, sd, prefix + "fusion_gate.0"))                           # :445-447
    if ctx is not None:
        g = ctx.apply(g, px.SITE_GATE, p)                                               # :448
    scores = _lin(g, sd, prefix + "fusion_gate.3").squeeze(-1)                          # :449,511
    scores = scores.masked_fill(invalid, -1e4)                                          # :513
    w = _softmax(scores, dim=-1)                                                        # :514
    w = torch.where(torch.isfinite(w), w, torch.zeros_like(w))                          # :516
    w = w / w.sum(dim=-1, keepdim=True).clamp_min(1e-8)                                 # :517-518
    fused = (w.unsqueeze(-1) * H_att).sum(dim=1)                                        # :521
    logits = _lin(fused, sd, prefix + "classifier")                                     # :522
    if return_info:
        return logits, {"fusion_weights": w, "per_modality_attention": attn_w, "H_att": H_att}
    return logits


def early_fusion(feats: Dict[str, torch.Tensor], mask, sd, prefix="fusion_model.",
                 modalities: Sequence[str] = ("audio", "video"), ctx: Optional[DropCtx] = None):
    """EarlyFusion.forward.  src/fusion.py:74-138: features multiplied by their mask column (no missing-modality
    embeddings exist in the reference, :124-128 falls to `x = m * x`), concatenated, then
    Linear-LN-ReLU-Dropout-Linear-LN-ReLU-Dropout-Linear (:60-72)."""
    parts = []
    for i, m in enumerate(modalities):
        x = feats[m]
        if mask is not None:
            x = torch.as_tensor(mask).to(x.dtype)[:, i:i + 1] * x
        parts.append(x)
    h = torch.cat(parts, dim=-1)
    p = ctx.p_fusion if ctx is not None else 0.0
    for lin, ln, site in (("0", "1", px.SITE_EARLY0), ("4", "5", px.SITE_EARLY1)):
        h = _lin(h, sd, f"{prefix}mlp.{lin}")
        h = torch.relu(_layer_norm(h, sd[f"{prefix}mlp.{ln}.weight"], sd[f"{prefix}mlp.{ln}.bias"]))
        if ctx is not None:
            h = ctx.apply(h, site, p)
    return _lin(h, sd, prefix + "mlp.8")


def late_fusion(feats: Dict[str, torch.Tensor], mask, sd, prefix="fusion_model.",
                modalities: Sequence[str] = ("audio", "video"), ctx: Optional[DropCtx] = None):
    """LateFusion.forward.  src/fusion.py:190-248: per-modality Linear-ReLU-Dropout-Linear classifiers (:180-188),
    weights softmax(fusion_logits) (:236), multiplied by the mask and renormalised with clamp_min(1e-8) (:238-241).
    Returns (fused_logits, [per-modality logits])."""
    p = ctx.p_fusion if ctx is not None else 0.0
    per = []
    for i, m in enumerate(modalities):
        h = torch.relu(_lin(feats[m], sd, f"{prefix}classifiers.{m}.0"))
        if ctx is not None:
            h = ctx.apply(h, px.SITE_LATE + i, p)
        per.append(_lin(h, sd, f"{prefix}classifiers.{m}.3"))
    stacked = torch.stack(per, dim=1)
    base_w = _softmax(sd[prefix + "fusion_logits"], dim=0)
    B = stacked.shape[0]
    if mask is not None:
        w = base_w.unsqueeze(0) * torch.as_tensor(mask).to(stacked.dtype)
        w = w / w.sum(dim=1, keepdim=True).clamp_min(1e-8)
    else:
        w = base_w.unsqueeze(0).expand(B, -1)
    return (w.unsqueeze(-1) * stacked).sum(dim=1), per


def late_fusion_with_uncertainty(feats: Dict[str, torch.Tensor], mask, sd, prefix="unc_model.",
                                 modalities: Sequence[str] = ("audio", "video"), ctx: Optional[DropCtx] = None,
                                 eps: float = 1e-6):
    """LateFusionWithUncertainty.forward (hidden_dim > 0).  src/fusion.py:629-641: per modality
    logits = classifiers[m](x) and u = softplus(uncert_heads[m](x)) with each head = Dropout, Linear, ReLU, Dropout,
    Linear (:613-620); a modality that is not given is zeros (:629).  UncertaintyWeightedFusion
    (src/uncertainty.py:303-317): w = mask / (u + eps), normalised by (sum + eps).
    Returns (fused (B, C), per-modality logits (B, M, C), weights (B, M))."""
    p = ctx.p_fusion if ctx is not None else 0.0
    B = torch.as_tensor(mask).shape[0]
    per, us = [], []
    for i, m in enumerate(modalities):
        x0 = feats[m] if m in feats else torch.zeros(B, sd[f"{prefix}classifiers.{m}.1.weight"].shape[1])
        outs = []
        for j, head in enumerate(("classifiers", "uncert_heads")):
            x = ctx.apply(x0, px.SITE_UNC + 4 * i + 2 * j, p) if ctx is not None else x0
            h = torch.relu(_lin(x, sd, f"{prefix}{head}.{m}.1"))
            if ctx is not None:
                h = ctx.apply(h, px.SITE_UNC + 4 * i + 2 * j + 1, p)
            outs.append(_lin(h, sd, f"{prefix}{head}.{m}.4"))
        per.append(outs[0])
        us.append(torch.nn.functional.softplus(outs[1]).squeeze(-1))
    logits = torch.stack(per, dim=1)
    uncert = torch.stack(us, dim=1)
    inv_w = (1.0 / (uncert + eps)) * torch.as_tensor(mask).to(logits.dtype)
    w = inv_w / (inv_w.sum(dim=1, keepdim=True) + eps)
    return (w.unsqueeze(-1) * logits).sum(dim=1), logits, w


def calibration_metrics(logits, labels, num_bins=15):
    """compute_calibration_metrics after the logits are collected (src/uncertainty.py:512-544) with
    CalibrationMetrics._bin_stats / ECE / MCE (:98-183): returns dict(ece, mce, nll, accuracy)."""
    logits = logits.float()
    labels = labels.long()
    n = labels.shape[0]
    logp = logits - torch.logsumexp(logits, dim=1, keepdim=True)
    nll = float(-(logp[torch.arange(n), labels]).sum() / max(1, n))
    probs = _softmax(logits, dim=1)
    conf, pred = probs.max(dim=1)
    conf = conf.clamp(0, 1)
    edges = torch.linspace(0, 1, steps=num_bins + 1)
    ece, mce, tot = 0.0, 0.0, 0
    stats = []
    for b in range(num_bins):
        lo, hi = edges[b], edges[b + 1]
        in_bin = (conf >= lo) & ((conf < hi) if b < num_bins - 1 else (conf <= hi))
        if in_bin.any():
            stats.append((int(in_bin.sum()), float(conf[in_bin].mean()), float((pred[in_bin] == labels[in_bin]).float().mean())))
    tot = sum(s[0] for s in stats)
    for nb, c, a in stats:
        ece += abs(a - c) * nb / tot
        mce = max(mce, abs(a - c))
    acc = float((pred == labels).float().mean()) if n else 0.0
    return {"ece": ece, "mce": mce, "nll": nll, "accuracy": acc}


def baseline_head(z_list, sd, prefix="fusion_head."):
    """Live path of MultimodalFusionModule.forward: concat -> Linear -> ReLU -> Linear.

    src/train.py:81-85 (ctor), :153-164 (forward; mask ignored).
    """
    fused = torch.cat(list(z_list), dim=-1)
    return _lin(torch.relu(_lin(fused, sd, prefix + "0")), sd, prefix + "2")


def av_forward(feats, frames, mask, sd, head="hybrid", ctx: Optional[DropCtx] = None,
               z_audio=None, hid=None):
    """encoders -> head, composed as src/train.py:133-136 + :138-147 (hybrid, the
    commented-out composition) or :153-164 (concat baseline, the live one).

    z_audio / hid let the caller pass the pass-invariant parts.
    """
    if z_audio is None:
        z_audio = sequence_encoder_lstm(feats, sd)
    if hid is None:
        hid = frame_hidden(frames, sd)
    z_video = frame_encoder_tail(hid, sd, ctx=ctx)
    if head == "hybrid":
        return hybrid_fusion({"audio": z_audio, "video": z_video}, mask, sd, ctx=ctx)
    if head == "early":
        return early_fusion({"audio": z_audio, "video": z_video}, mask, sd, prefix="early_model.", ctx=ctx)
    if head == "late":
        return late_fusion({"audio": z_audio, "video": z_video}, mask, sd, prefix="late_model.", ctx=ctx)[0]
    return baseline_head([z_audio, z_video], sd)


def mc_statistics(logits_stack):
    """MCDropoutUncertainty.forward after the T passes.  src/uncertainty.py:74-84:
    mean of LOGITS over passes; POPULATION variance of softmax PROBABILITIES over
    passes, averaged over classes."""
    mean_logits = logits_stack.mean(dim=0)
    probs = _softmax(logits_stack, dim=-1)
    var = ((probs - probs.mean(dim=0, keepdim=True)) ** 2).mean(dim=0)
    return mean_logits, var.mean(dim=-1)


def mc_dropout_forward_as_reference(feats, frames, mask, sd, T, seed, clip_ids, head="hybrid",
                                    p_encoder=0.1, p_fusion=0.3):
    """MCDropoutUncertainty exactly as the reference executes it (src/uncertainty.py:62-65): T FULL forwards of
    the model - audio encoder (torch.nn.LSTM) and video Linear + ReLU recomputed every pass.  Same values as
    mc_dropout_forward; this is the one the CPU baseline times."""
    stack = []
    with torch.inference_mode():
        for t in range(T):
            ctx = DropCtx(seed=seed, t=t, clip_ids=np.asarray(clip_ids), p_encoder=p_encoder, p_fusion=p_fusion)
            z_audio = sequence_encoder_lstm_torch(feats, sd)
            stack.append(av_forward(None, frames, mask, sd, head, ctx, z_audio=z_audio))
        stack = torch.stack(stack, dim=0)
        mean_logits, unc = mc_statistics(stack)
    return mean_logits, unc, stack


def mc_dropout_forward(feats, frames, mask, sd, T, seed, clip_ids, head="hybrid",
                       p_encoder=0.1, p_fusion=0.3):
    """MCDropoutUncertainty(model, T).forward(features, mask).  src/uncertainty.py:48-86.

    The reference recomputes the whole model T times; the audio encoder and the
    video Linear+ReLU contain no active nn.Dropout, so they are
    computed once here - same values, T times less CPU work.
    Returns (mean_logits (B,C), uncertainty (B,), logits_stack (T,B,C)).
    """
    z_audio = sequence_encoder_lstm(feats, sd)
    hid = frame_hidden(frames, sd)
    stack = []
    for t in range(T):
        ctx = DropCtx(seed=seed, t=t, clip_ids=np.asarray(clip_ids), p_encoder=p_encoder,
                      p_fusion=p_fusion)
        stack.append(av_forward(None, None, mask, sd, head, ctx, z_audio=z_audio, hid=hid))
    stack = torch.stack(stack, dim=0)
    mean_logits, unc = mc_statistics(stack)
    return mean_logits, unc, stack


def to_torch(sd_np, dtype=torch.float32):
    return {k: torch.from_numpy(np.asarray(v)).to(dtype) for k, v in sd_np.items()}
