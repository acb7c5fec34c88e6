#!/usr/bin/env python
"""Golden vectors for the SURVEY.md 8(f) sequence-attention modules, from the REFERENCE's own classes.

Runs only in the build container (needs /root/reference).  Instantiates, with torch's default init under a fixed
seed, the reference's

    SequenceEncoder(encoder_type='transformer')     src/encoders.py:96-113, :170-192
    TemporalAttention (+ pool_sequence)             src/attention.py:143-274
    PairwiseModalityAttention                       src/attention.py:421-548 (second definition)

runs them on seeded inputs (eval, and stochastic passes with the Philox spec masks of oracle/philox.py injected into
their nn.Dropout modules), checks oracle/model.py's restatements against them, and writes the weights (keys "w.<prefix><state-dict key>"), inputs and
outputs to tests/golden/seq_golden.npz (small dimensions, so the fixture stays small; the GPU tests also compare the
CUDA path with the pinned oracle at the benchmark dimensions).

    python tests/golden/make_golden_seq.py
"""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")
sys.dont_write_bytecode = True

import numpy as np
import torch
import torch.nn as nn

import attention as ref_attention        # /root/reference/src/attention.py
import encoders as ref_encoders          # /root/reference/src/encoders.py

from oracle import model as om, philox as px

SEED = 42
CLIP_IDS = np.array([0, 1, 1000003, 4294967295], dtype=np.uint64)
POS_ROWS = 64                             # rows of pos_embedding kept in the fixture (the rest are zeroed)


def spec_mask(x, clip_axis, t, site, p, clip_ids=CLIP_IDS):
    """x * keep / (1 - p) with the spec mask; the per-clip mask is x's shape without the clip axis, row-major."""
    xm = x.movedim(clip_axis, 0)
    n = int(np.prod(xm.shape[1:]))
    keep = px.dropout_keep(n, clip_ids, t, site, p, SEED).reshape(xm.shape)
    scaled = xm * (torch.from_numpy(keep).float() * float(px.keep_scale(p)))
    return scaled.movedim(0, clip_axis)


def main():
    warnings.filterwarnings("ignore")
    torch.set_num_threads(8)
    out = {"clip_ids": CLIP_IDS, "seed": np.int64(SEED)}
    worst = {}

    def rec(name, ref_t, ora_t):
        ref_t = ref_t.detach()
        out[name] = ref_t.numpy()
        a, b = ref_t.double(), ora_t.double()
        same_nan = torch.isnan(a) == torch.isnan(b)
        assert bool(same_nan.all()), name
        d = (torch.nan_to_num(a) - torch.nan_to_num(b)).abs().max().item()
        worst[name] = (d, torch.nan_to_num(a).abs().max().item())

    # ---- SequenceEncoder(encoder_type='transformer') ---------------------------------------------------------------
    torch.manual_seed(7)
    B = len(CLIP_IDS)
    tfm = ref_encoders.SequenceEncoder(input_dim=104, hidden_dim=64, output_dim=32, num_layers=2,
                                       encoder_type="transformer", dropout=0.1)
    with torch.no_grad():
        tfm.pos_embedding.weight[POS_ROWS:].zero_()
    tfm.eval()
    tsd = {"tfm_enc." + k: v.detach().clone() for k, v in tfm.state_dict().items()}
    for k, v in tsd.items():
        out["w." + k] = v.numpy()[:POS_ROWS] if k.endswith("pos_embedding.weight") else v.numpy()
    x_long = torch.randn(B, 37, 104)
    x_short = torch.randn(B, 9, 104)          # T * H = 576 < 1024: the 32-bit-draw form of the H-wide sites
    lengths = torch.tensor([37, 5, 1, 20])
    out["tfm.x_long"], out["tfm.x_short"], out["tfm.lengths"] = x_long.numpy(), x_short.numpy(), lengths.numpy()
    with torch.inference_mode():
        rec("tfm.z_long", tfm(x_long), om.sequence_encoder_transformer(x_long, tsd))
        rec("tfm.z_long_lengths", tfm(x_long, lengths), om.sequence_encoder_transformer(x_long, tsd, lengths=lengths))
        rec("tfm.z_short", tfm(x_short), om.sequence_encoder_transformer(x_short, tsd))
    state = {"t": 0}
    found = 0
    for name, mod in tfm.named_modules():
        if isinstance(mod, nn.Dropout):
            parts = name.split(".")                 # transformer.layers.<l>.<dropout1|dropout|dropout2>
            site = px.SITE_TFM + 3 * int(parts[2]) + {"dropout1": 0, "dropout": 1, "dropout2": 2}[parts[3]]
            mod.forward = (lambda x, mod=mod, site=site: spec_mask(x, 1, state["t"], site, mod.p))   # x is (T, B, .)
            mod.train()                              # what MCDropoutUncertainty._set_dropout_train does
            found += 1
    assert found == 6, found
    with torch.inference_mode():
        for t in range(2):
            state["t"] = t
            ctx = om.DropCtx(SEED, t, CLIP_IDS)
            rec(f"tfm.z_long_pass{t}", tfm(x_long, lengths), om.sequence_encoder_transformer(x_long, tsd, lengths=lengths, ctx=ctx))
            rec(f"tfm.z_short_pass{t}", tfm(x_short), om.sequence_encoder_transformer(x_short, tsd, ctx=ctx))

    # ---- TemporalAttention ----------------------------------------------------------------------------------------------
    torch.manual_seed(8)
    ta = ref_attention.TemporalAttention(feature_dim=40, hidden_dim=64, num_heads=4, dropout=0.1).eval()
    asd = {"tattn." + k: v.detach().clone() for k, v in ta.state_dict().items()}
    for k, v in asd.items():
        out["w." + k] = v.numpy()
    seq = torch.randn(B, 19, 40)
    tmask = torch.ones(B, 19)
    tmask[1, 11:] = 0
    tmask[2, ::2] = 0
    tmask[3, 1:] = 0
    out["tattn.seq"], out["tattn.mask"] = seq.numpy(), tmask.numpy()
    with torch.inference_mode():
        for tag, mk in (("nomask", None), ("mask", tmask)):
            att, w = ta(seq, mk)
            oatt, ow = om.temporal_attention(seq, mk, asd)
            rec(f"tattn.attended_{tag}", att, oatt)
            rec(f"tattn.weights_{tag}", w, ow)
            rec(f"tattn.pooled_{tag}", ta.pool_sequence(att, w), om.temporal_pool_sequence(oatt, ow))
        ta.train()                                   # the module's dropouts are dead code: train mode changes nothing
        att_tr, _ = ta(seq, tmask)
        assert torch.equal(att_tr, ta.eval()(seq, tmask)[0])

    # ---- PairwiseModalityAttention ----------------------------------------------------------------------------------------
    torch.manual_seed(9)
    dims = {"audio": 128, "video": 128, "imu": 64}
    pm = ref_attention.PairwiseModalityAttention(dims, hidden_dim=64, num_heads=4, dropout=0.1).eval()
    psd = {"pair." + k: v.detach().clone() for k, v in pm.state_dict().items()}
    for k, v in psd.items():
        out["w." + k] = v.numpy()
    pids = np.array([0, 1, 2, 3, 1000003, 4294967295], dtype=np.uint64)
    out["pair.clip_ids"] = pids
    pf = {m: torch.randn(len(pids), d) for m, d in dims.items()}
    for m in dims:
        out[f"pair.x_{m}"] = pf[m].numpy()
    pmasks = {"mixed": torch.tensor([[1, 1, 1], [1, 0, 1], [0, 1, 1], [0, 0, 1], [1, 1, 0], [0, 0, 0]], dtype=torch.float32),
              "noimu": torch.tensor([[1, 1, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [1, 1, 0], [0, 0, 0]], dtype=torch.float32)}
    with torch.inference_mode():
        for tag, mk in (("nomask", None), ("mixed", pmasks["mixed"]), ("noimu", pmasks["noimu"])):
            if mk is not None:
                out[f"pair.mask_{tag}"] = mk.numpy()
            att, maps = pm(pf, None if mk is None else mk.clone())
            oatt, omaps = om.pairwise_modality_attention(pf, mk, psd, dims)
            assert sorted(maps) == sorted(omaps), (sorted(maps), sorted(omaps))
            out[f"pair.map_keys_{tag}"] = np.array(sorted(maps))
            for m in dims:
                rec(f"pair.out_{tag}_{m}", att[m], oatt[m])
            for k in maps:
                rec(f"pair.map_{tag}_{k}", maps[k], omaps[k])
    # one stochastic pass (t = 3): attention dropout of every direction + msg_dropout (one module, called per modality)
    T_PASS = 3
    order = list(pm.cross_attn.keys())
    for key in order:
        mod = pm.cross_attn[key].attn_dropout
        mod.forward = (lambda x, mod=mod, site=px.SITE_ATTN + order.index(key): spec_mask(x, 0, T_PASS, site, mod.p, pids))
        mod.train()
    call = {"n": 0}

    def msg_drop(x, mod=pm.msg_dropout):
        site = px.SITE_PAIR_MSG + call["n"] % len(dims)
        call["n"] += 1
        return spec_mask(x, 0, T_PASS, site, mod.p, pids)
    pm.msg_dropout.forward = msg_drop
    pm.msg_dropout.train()
    with torch.inference_mode():
        att, maps = pm(pf, pmasks["mixed"].clone())
        oatt, omaps = om.pairwise_modality_attention(pf, pmasks["mixed"], psd, dims, ctx=om.DropCtx(SEED, T_PASS, pids))
        for m in dims:
            rec(f"pair.out_drop_{m}", att[m], oatt[m])
        for k in maps:
            rec(f"pair.map_drop_{k}", maps[k], omaps[k])
    out["pair.t_pass"] = np.int64(T_PASS)

    print("reference vs oracle/model.py (max abs diff, max |ref|):")
    bad = False
    for k_, (d, s) in worst.items():
        flag = "" if d <= 2e-5 * max(1.0, s) else "   <-- MISMATCH"
        bad |= bool(flag)
        print(f"  {k_:40s} {d:.3e}  {s:.3e}{flag}")
    assert not bad
    path = os.path.join(HERE, "seq_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
