"""oracle/model.py against the golden vectors produced by the REFERENCE's own modules
(tests/golden/make_golden.py imported them from /root/reference/src in the build container)."""
import numpy as np
import torch

from oracle import frontend as fe, model as om, philox as px, synth

TOL = dict(rtol=1e-4, atol=2e-5)


def _inputs(golden):
    ids = golden["clip_ids"]
    seed = int(golden["seed"])
    sd = om.to_torch(synth.synth_weights(seed))
    wave = synth.synth_audio(ids, seed)
    frames = torch.from_numpy(synth.synth_video(ids, seed))
    feats = torch.from_numpy(fe.logmel_mfcc(wave))
    masks = {
        "ones": torch.ones(len(ids), 2),
        "mixed": torch.tensor([[1, 1], [1, 0], [0, 1], [0, 0], [1, 1], [0, 1]], dtype=torch.float32),
        "p05": torch.from_numpy(px.modality_mask(ids, 2, 0.05, seed)),
    }
    return ids, seed, sd, feats, frames, masks


def test_encoders_and_heads(golden):
    torch.set_num_threads(8)
    ids, seed, sd, feats, frames, masks = _inputs(golden)
    z_a = om.sequence_encoder_lstm(feats, sd)
    z_v = om.frame_encoder(frames, sd)
    np.testing.assert_allclose(z_a.numpy(), golden["z_audio"], **TOL)
    np.testing.assert_allclose(z_v.numpy(), golden["z_video"], **TOL)
    hid = om.frame_hidden(frames, sd)
    for name, mask in masks.items():
        lg, info = om.hybrid_fusion({"audio": z_a, "video": z_v}, mask, sd, return_info=True)
        np.testing.assert_allclose(lg.numpy(), golden[f"logits_hybrid_{name}"], **TOL)
        np.testing.assert_allclose(info["fusion_weights"].numpy(), golden[f"fusion_weights_{name}"], **TOL)
        np.testing.assert_allclose(info["per_modality_attention"][0].numpy(), golden[f"attn_audio_{name}"], **TOL)
        np.testing.assert_allclose(info["per_modality_attention"][1].numpy(), golden[f"attn_video_{name}"], **TOL)
    lg = om.av_forward(None, None, None, sd, "concat", z_audio=z_a, hid=hid)
    np.testing.assert_allclose(lg.numpy(), golden["logits_concat_ones"], **TOL)
    # MC-dropout through the oracle's Philox masks == the reference wrapper over mask-injected reference modules
    T = int(golden["t_mc"])
    for head, names in (("hybrid", ("ones", "p05")), ("concat", ("ones",)), ("early", ("ones", "p05")),
                        ("late", ("ones", "p05"))):
        for name in names:
            mean, unc, _ = om.mc_dropout_forward(feats, frames, masks[name], sd, T, seed, ids, head)
            np.testing.assert_allclose(mean.numpy(), golden[f"mc_mean_logits_{head}_{name}"], **TOL)
            np.testing.assert_allclose(unc.numpy(), golden[f"mc_uncertainty_{head}_{name}"], rtol=1e-3, atol=1e-7)
            assert (mean.argmax(dim=1).numpy() == golden[f"mc_labels_{head}_{name}"]).all()


def test_early_late_fusion_and_calibration(golden):
    """EarlyFusion / LateFusion / CalibrationMetrics restatements against the reference's
    own outputs (make_golden.py)."""
    ids, seed, sd, feats, frames, masks = _inputs(golden)
    z = {"audio": om.sequence_encoder_lstm(feats, sd), "video": om.frame_encoder(frames, sd)}
    for name, mask in masks.items():
        lg = om.early_fusion(z, mask, sd, prefix="early_model.")
        np.testing.assert_allclose(lg.numpy(), golden[f"logits_early_{name}"], **TOL)
        lg, per = om.late_fusion(z, mask, sd, prefix="late_model.")
        np.testing.assert_allclose(lg.numpy(), golden[f"logits_late_{name}"], **TOL)
        if name == "mixed":
            np.testing.assert_allclose(per[0].numpy(), golden["late_per_audio_mixed"], **TOL)
            np.testing.assert_allclose(per[1].numpy(), golden["late_per_video_mixed"], **TOL)
    cal = om.calibration_metrics(torch.from_numpy(golden["cal.logits"]), torch.from_numpy(golden["cal.labels"]), 15)
    for k in ("ece", "mce", "nll", "accuracy"):
        assert abs(cal[k] - float(golden["cal." + k])) <= 1e-6, k
    cal10 = om.calibration_metrics(torch.from_numpy(golden["cal.logits"]), torch.from_numpy(golden["cal.labels"]), 10)
    assert abs(cal10["ece"] - float(golden["cal.ece10"])) <= 1e-6


def test_gru_sequence_encoder(golden):
    """SequenceEncoder(encoder_type='gru') restatement against the reference module's outputs."""
    ids, seed, sd, feats, frames, masks = _inputs(golden)
    np.testing.assert_allclose(om.sequence_encoder_gru(feats, sd).numpy(), golden["gru.z"], **TOL)
    z = om.sequence_encoder_gru(feats, sd, lengths=torch.from_numpy(golden["gru.lengths"]))
    np.testing.assert_allclose(z.numpy(), golden["gru.z_lengths"], **TOL)


def test_cnn_sequence_encoder(golden):
    """SequenceEncoder(encoder_type='cnn') restatement against the reference module (eval and
    three stochastic passes with the spec masks injected into its twice-used nn.Dropout)."""
    ids, seed, sd, feats, frames, masks = _inputs(golden)
    np.testing.assert_allclose(om.sequence_encoder_cnn(feats, sd).numpy(), golden["cnn.z"], rtol=1e-4, atol=1e-4)
    for t in range(3):
        z = om.sequence_encoder_cnn(feats, sd, ctx=om.DropCtx(seed, t, ids))
        np.testing.assert_allclose(z.numpy(), golden[f"cnn.z_pass{t}"], rtol=1e-4, atol=1e-4)


def test_late_fusion_with_uncertainty(golden):
    """LateFusionWithUncertainty + UncertaintyWeightedFusion restatement vs the reference."""
    ids, seed, sd, feats, frames, masks = _inputs(golden)
    z = {"audio": torch.from_numpy(golden["z_audio"]), "video": torch.from_numpy(golden["z_video"])}
    for name in ("ones", "mixed"):
        lg, per, w = om.late_fusion_with_uncertainty(z, masks[name], sd)
        np.testing.assert_allclose(lg.numpy(), golden[f"unc.logits_{name}"], **TOL)
        np.testing.assert_allclose(per.numpy(), golden[f"unc.per_{name}"], **TOL)
        np.testing.assert_allclose(w.numpy(), golden[f"unc.weights_{name}"], **TOL)
    lg, _, _ = om.late_fusion_with_uncertainty({"video": z["video"]}, masks["mixed"], sd)
    np.testing.assert_allclose(lg.numpy(), golden["unc.logits_video_only"], **TOL)
    lg, _, w = om.late_fusion_with_uncertainty(z, masks["mixed"], sd, ctx=om.DropCtx(seed, 2, ids))
    np.testing.assert_allclose(lg.numpy(), golden["unc.logits_drop"], **TOL)
    np.testing.assert_allclose(w.numpy(), golden["unc.weights_drop"], **TOL)


def test_cross_modal_attention_general_shapes(golden):
    sd = {k[len("xattn."):]: torch.from_numpy(golden[k]) for k in golden.files if k.startswith("xattn.att.")}
    q, kv = torch.from_numpy(golden["xattn.q"]), torch.from_numpy(golden["xattn.kv"])
    inv = torch.from_numpy(golden["xattn.invalid"])
    out, attn = om.cross_modal_attention(q, kv, kv, inv, sd, "att.", 4)
    np.testing.assert_allclose(out.numpy(), golden["xattn.out"], **TOL)
    np.testing.assert_allclose(attn.numpy(), golden["xattn.attn"], **TOL)
    assert (attn[2] == 0).all()                              # all keys masked -> zeros (src/attention.py:124-128)


def test_fp64_oracle_close_to_fp32(golden):
    ids, seed, sd, feats, frames, masks = _inputs(golden)
    sd64 = {k: v.double() for k, v in sd.items()}
    z32 = om.frame_encoder(frames[:2], sd)
    z64 = om.frame_encoder(frames[:2].double(), sd64)
    assert (z32.double() - z64).abs().max() < 1e-5


def test_reference_style_execution_matches_hoisted(golden):
    """The CPU-baseline path (torch.nn.LSTM, T full forwards like src/uncertainty.py:62-65) computes the same
    numbers as the hoisted, written-out oracle."""
    ids, seed, sd, feats, frames, masks = _inputs(golden)
    z1 = om.sequence_encoder_lstm(feats[:3], sd)
    z2 = om.sequence_encoder_lstm_torch(feats[:3], sd)
    np.testing.assert_allclose(z2.numpy(), z1.numpy(), rtol=1e-4, atol=2e-6)
    a = om.mc_dropout_forward(feats[:3], frames[:3], masks["p05"][:3], sd, 3, seed, ids[:3])
    b = om.mc_dropout_forward_as_reference(feats[:3], frames[:3], masks["p05"][:3], sd, 3, seed, ids[:3])
    np.testing.assert_allclose(b[0].numpy(), a[0].numpy(), rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(b[1].numpy(), a[1].numpy(), rtol=1e-3, atol=1e-8)
