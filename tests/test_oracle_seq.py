"""oracle/model.py restatements of the SURVEY.md 8(f) sequence-attention modules against the golden vectors the
REFERENCE's own classes produced (tests/golden/make_golden_seq.py), and the state-dict layouts of the drop-in modules."""
import numpy as np
import torch

from oracle import model as om

from _seq_fixture import PAIR_DIMS, load, state_dict

TOL = dict(rtol=1e-4, atol=2e-5)


def test_transformer_encoder_oracle_vs_reference_golden():
    g = load()
    sd = state_dict(g, "tfm_enc.", keep_prefix=True)
    ids, seed = g["clip_ids"], int(g["seed"])
    xl, xs, ln = torch.from_numpy(g["tfm.x_long"]), torch.from_numpy(g["tfm.x_short"]), torch.from_numpy(g["tfm.lengths"])
    np.testing.assert_allclose(om.sequence_encoder_transformer(xl, sd).numpy(), g["tfm.z_long"], **TOL)
    np.testing.assert_allclose(om.sequence_encoder_transformer(xl, sd, lengths=ln).numpy(), g["tfm.z_long_lengths"], **TOL)
    np.testing.assert_allclose(om.sequence_encoder_transformer(xs, sd).numpy(), g["tfm.z_short"], **TOL)
    for t in range(2):
        ctx = om.DropCtx(seed, t, ids)
        np.testing.assert_allclose(om.sequence_encoder_transformer(xl, sd, lengths=ln, ctx=ctx).numpy(),
                                   g[f"tfm.z_long_pass{t}"], **TOL)
        np.testing.assert_allclose(om.sequence_encoder_transformer(xs, sd, ctx=ctx).numpy(), g[f"tfm.z_short_pass{t}"], **TOL)


def test_temporal_attention_oracle_vs_reference_golden():
    g = load()
    sd = state_dict(g, "tattn.", keep_prefix=True)
    seq, mask = torch.from_numpy(g["tattn.seq"]), torch.from_numpy(g["tattn.mask"])
    for tag, mk in (("nomask", None), ("mask", mask)):
        att, w = om.temporal_attention(seq, mk, sd)
        np.testing.assert_allclose(att.numpy(), g[f"tattn.attended_{tag}"], **TOL)
        np.testing.assert_allclose(w.numpy(), g[f"tattn.weights_{tag}"], **TOL)
        np.testing.assert_allclose(om.temporal_pool_sequence(att, w).numpy(), g[f"tattn.pooled_{tag}"], **TOL)


def test_pairwise_attention_oracle_vs_reference_golden():
    g = load()
    sd = state_dict(g, "pair.", keep_prefix=True)
    feats = {m: torch.from_numpy(g[f"pair.x_{m}"]) for m in PAIR_DIMS}
    for tag in ("nomask", "mixed", "noimu"):
        mk = None if tag == "nomask" else torch.from_numpy(g[f"pair.mask_{tag}"])
        out, maps = om.pairwise_modality_attention(feats, mk, sd, PAIR_DIMS)
        assert sorted(maps) == list(g[f"pair.map_keys_{tag}"])
        for m in PAIR_DIMS:
            np.testing.assert_allclose(out[m].numpy(), g[f"pair.out_{tag}_{m}"], **TOL)
        for k in maps:
            np.testing.assert_allclose(maps[k].numpy(), g[f"pair.map_{tag}_{k}"], **TOL)
    ctx = om.DropCtx(int(g["seed"]), int(g["pair.t_pass"]), g["pair.clip_ids"])
    out, maps = om.pairwise_modality_attention(feats, torch.from_numpy(g["pair.mask_mixed"]), sd, PAIR_DIMS, ctx=ctx)
    for m in PAIR_DIMS:
        np.testing.assert_allclose(out[m].numpy(), g[f"pair.out_drop_{m}"], **TOL)
    for k in maps:
        np.testing.assert_allclose(maps[k].numpy(), g[f"pair.map_drop_{k}"], **TOL)


def test_drop_in_modules_load_the_reference_state_dicts():
    """The reference classes' state-dicts (keys and shapes as stored in the fixture) load strict=True (no GPU needed)."""
    import multimodal_emotion_detection_b200 as pkg
    g = load()
    enc = pkg.SequenceEncoder(input_dim=104, hidden_dim=64, output_dim=32, num_layers=2, encoder_type="transformer")
    enc.load_state_dict(state_dict(g, "tfm_enc."), strict=True)
    ta = pkg.TemporalAttention(feature_dim=40, hidden_dim=64, num_heads=4)
    ta.load_state_dict(state_dict(g, "tattn."), strict=True)
    pm = pkg.PairwiseModalityAttention(PAIR_DIMS, hidden_dim=64, num_heads=4)
    pm.load_state_dict(state_dict(g, "pair."), strict=True)
    assert [m.site_index for m in pm.cross_attn.values()] == list(range(6))
    assert list(pm.cross_attn.keys()) == ["audio_to_video", "video_to_audio", "audio_to_imu", "imu_to_audio",
                                         "video_to_imu", "imu_to_video"]
    x = torch.zeros(2, 5, 104)
    try:
        enc(x)
    except RuntimeError as e:                          # CPU tensors: no fallback
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("CPU tensors must be rejected")
