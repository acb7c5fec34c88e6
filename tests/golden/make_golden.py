#!/usr/bin/env python
"""Generate the golden vectors in tests/golden/ by running the REFERENCE's own modules.

Runs only in the build container, where /root/reference exists (it cannot travel
to the GPU box).  Imports encoders / attention / fusion / uncertainty from
/root/reference/src (flat imports, like src/train.py:28-31), composes them the
way MultimodalFusionModule does (src/train.py:41-85 ctor, :121-164 forward, and
the commented hybrid composition :138-147), loads the seeded synthetic weights
with strict=True (proving the state-dict layout of SURVEY.md Appendix A), and
records outputs on seeded synthetic inputs.

Dropout: every nn.Dropout instance of the reference model is given a forward that
multiplies by the Philox spec mask (oracle/philox.py) instead of torch's global
mt19937 stream; the reference's MCDropoutUncertainty then runs UNMODIFIED on top.

It also checks oracle/model.py (the restatement that travels) against the
reference modules and prints the max differences.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")
sys.dont_write_bytecode = True

import numpy as np
import torch
import torch.nn as nn

import encoders as ref_encoders          # /root/reference/src/encoders.py
import fusion as ref_fusion              # /root/reference/src/fusion.py
import uncertainty as ref_uncertainty    # /root/reference/src/uncertainty.py

from oracle import frontend as ofe, model as om, philox as px, synth

SEED = 42
CLIP_IDS = np.array([0, 1, 2, 3, 1000003, 4294967295], dtype=np.uint64)
T_MC = 8
MODALITIES = ["audio", "video"]
ENC_CFG = {   # configs/base.yaml:30-47 with the BASELINE.json input dims
    "audio": {"type": "sequence", "input_dim": 104, "encoder_type": "lstm", "hidden_dim": 256,
              "output_dim": 128, "num_layers": 2, "dropout": 0.1},
    "video": {"type": "frame", "input_dim": 12544, "temporal_pooling": "attention",
              "hidden_dim": 256, "output_dim": 128, "dropout": 0.1},
}


class RefAVModel(nn.Module):
    """MultimodalFusionModule without Lightning: src/train.py:50-85 and :121-164."""

    def __init__(self, head):
        super().__init__()
        self.head = head
        self.encoders = nn.ModuleDict()
        dims = {}
        for m in MODALITIES:                                   # train.py:55-66
            cfg = ENC_CFG[m]
            self.encoders[m] = ref_encoders.build_encoder(
                modality=m, input_dim=cfg.get("input_dim", 64), output_dim=128, encoder_config=cfg)
            dims[m] = 128
        self.fusion_model = ref_fusion.build_fusion_model(     # train.py:69-76 (commented there)
            fusion_type=head if head in ("early", "late") else "hybrid", modality_dims=dims, num_classes=8,
            hidden_dim=256, num_heads=4, dropout=0.3)
        self.fusion_head = nn.Sequential(                      # train.py:81-85
            nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, 8))

    def forward(self, features, mask=None):
        enc = {m: e(features[m]) for m, e in self.encoders.items() if m in features}  # :134-136
        if self.head in ("hybrid", "early", "late"):
            out = self.fusion_model(enc, mask)                 # :139
            return out[0] if isinstance(out, tuple) else out   # :142-145
        fused = torch.cat([enc[m] for m in MODALITIES if m in enc], dim=-1)           # :153-161
        return self.fusion_head(fused)                         # :163


class MaskInjector:
    """Gives each nn.Dropout of the reference model its Philox spec mask."""

    ENC_SITES = {
        "encoders.video.frame_mlp.2": px.SITE_FRAME_MLP,
        "encoders.video._dropout": px.SITE_FRAME_POOL,
    }
    HEAD_SITES = {
        "hybrid": {"fusion_model.attn_layers.audio.attn_dropout": px.SITE_ATTN + 0,
                   "fusion_model.attn_layers.video.attn_dropout": px.SITE_ATTN + 1,
                   "fusion_model.fusion_gate.2": px.SITE_GATE},
        "early": {"fusion_model.mlp.3": px.SITE_EARLY0, "fusion_model.mlp.7": px.SITE_EARLY1},
        "late": {"fusion_model.classifiers.audio.2": px.SITE_LATE + 0,
                 "fusion_model.classifiers.video.2": px.SITE_LATE + 1},
    }

    def __init__(self, model, seed, clip_ids):
        self.seed, self.clip_ids, self.t = seed, clip_ids, -1
        self.SITES = dict(self.ENC_SITES)
        self.SITES.update(self.HEAD_SITES["hybrid" if model.head == "concat" else model.head])
        model.register_forward_pre_hook(self._bump)
        found = {n for n, m in model.named_modules() if isinstance(m, nn.Dropout)}
        assert found == set(self.SITES), (found, set(self.SITES))
        for name, mod in model.named_modules():
            if isinstance(mod, nn.Dropout):
                mod.forward = self._make_forward(mod, self.SITES[name])

    def _bump(self, *_):
        self.t += 1

    def _make_forward(self, mod, site):
        def fwd(x):
            if not mod.training or mod.p <= 0:
                return x
            n = int(np.prod(x.shape[1:]))
            keep = px.dropout_keep(n, self.clip_ids, self.t, site, mod.p, self.seed).reshape(x.shape)
            return x * (torch.from_numpy(keep).float() * float(px.keep_scale(mod.p)))
        return fwd


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    sd_np = synth.synth_weights(SEED)
    sd = om.to_torch(sd_np)
    wave = synth.synth_audio(CLIP_IDS, SEED)
    frames_u8 = synth.synth_video_u8(CLIP_IDS, SEED)
    frames = torch.from_numpy(frames_u8.astype(np.float32) / np.float32(255.0))
    feats_np = ofe.logmel_mfcc(wave)                       # oracle front-end (no reference exists)
    feats = torch.from_numpy(feats_np)
    B = len(CLIP_IDS)
    masks = {
        "ones": torch.ones(B, 2),
        "mixed": torch.tensor([[1, 1], [1, 0], [0, 1], [0, 0], [1, 1], [0, 1]], dtype=torch.float32),
        "p05": torch.from_numpy(px.modality_mask(CLIP_IDS, 2, 0.05, SEED)),
    }

    out = {"clip_ids": CLIP_IDS, "seed": np.int64(SEED), "t_mc": np.int64(T_MC)}
    # front-end anchors (oracle's own output, stored so drift in the oracle is caught)
    out["feats_clip0"] = feats_np[0]
    out["feats_last"] = feats_np[-1]
    worst = {}

    def rec(name, ref_t, ora_t):
        ref_t = ref_t.detach()
        out[name] = ref_t.numpy()
        d = (ref_t.double() - ora_t.double()).abs().max().item()
        s = ref_t.abs().max().item()
        worst[name] = (d, s)

    for head in ("hybrid", "concat", "early", "late"):
        model = RefAVModel(head)
        full = dict(synth.state_dict_for("hybrid" if head == "concat" else head, sd))
        full.update({k: v for k, v in sd.items() if k.startswith("fusion_head.")})
        missing = model.load_state_dict(full, strict=True)       # layout check, Appendix A
        model.eval()
        inj = MaskInjector(model, SEED, CLIP_IDS)
        with torch.inference_mode():
            z_a = model.encoders["audio"](feats)
            z_v = model.encoders["video"](frames)
            if head == "hybrid":
                rec("z_audio", z_a, om.sequence_encoder_lstm(feats, sd))
                rec("z_video", z_v, om.frame_encoder(frames, sd))
            for mname, mask in masks.items():
                if head == "concat" and mname != "ones":
                    continue                                     # mask ignored on the live path
                logits = model({"audio": feats, "video": frames}, mask.clone())
                rec(f"logits_{head}_{mname}", logits,
                    om.av_forward(feats, frames, mask, sd, head))
                if head == "hybrid":
                    lg, info = model.fusion_model({"audio": z_a, "video": z_v}, mask.clone(),
                                                  return_attention=True)
                    _, oinfo = om.hybrid_fusion({"audio": z_a, "video": z_v}, mask, sd,
                                                return_info=True)
                    rec(f"fusion_weights_{mname}", info["fusion_weights"], oinfo["fusion_weights"])
                    rec(f"attn_audio_{mname}", info["per_modality_attention"]["audio"],
                        oinfo["per_modality_attention"][0])
                    rec(f"attn_video_{mname}", info["per_modality_attention"]["video"],
                        oinfo["per_modality_attention"][1])
        # MC-dropout: the reference wrapper, unmodified, over the mask-injected model
        if head in ("early", "late"):
            with torch.inference_mode():
                for mname in ("mixed",):
                    logits = model({"audio": feats, "video": frames}, masks[mname].clone())
                    rec(f"logits_{head}_{mname}", logits, om.av_forward(feats, frames, masks[mname], sd, head))
                if head == "late":
                    _, per = model.fusion_model({"audio": z_a, "video": z_v}, masks["mixed"].clone())
                    _, oper = om.late_fusion({"audio": z_a, "video": z_v}, masks["mixed"], sd, prefix="late_model.")
                    rec("late_per_audio_mixed", per["audio"], oper[0])
                    rec("late_per_video_mixed", per["video"], oper[1])
        for mname in (("ones", "p05") if head != "concat" else ("ones",)):
            mask = masks[mname]
            inj.t = -1
            mc = ref_uncertainty.MCDropoutUncertainty(model, num_samples=T_MC)
            mean_logits, unc = mc({"audio": feats, "video": frames}, mask.clone())
            o_mean, o_unc, o_stack = om.mc_dropout_forward(feats, frames, mask, sd, T_MC, SEED,
                                                           CLIP_IDS, head)
            rec(f"mc_mean_logits_{head}_{mname}", mean_logits, o_mean)
            rec(f"mc_uncertainty_{head}_{mname}", unc, o_unc)
            out[f"mc_labels_{head}_{mname}"] = mean_logits.argmax(dim=1).numpy()
            assert (mean_logits.argmax(dim=1) == o_mean.argmax(dim=1)).all()

    # standalone CrossModalAttention at the reference's own smoke shapes
    # (src/attention.py:605-651: (4,512) x (4,64) -> (4,256)) with a sequence of keys
    torch.manual_seed(1)
    import attention as ref_attention
    att = ref_attention.CrossModalAttention(query_dim=512, key_dim=64, hidden_dim=256, num_heads=4,
                                            dropout=0.1).eval()
    asd = {"att." + k: v for k, v in att.state_dict().items()}
    q = torch.randn(4, 3, 512)
    kv = torch.randn(4, 5, 64)
    inv = torch.tensor([[0, 0, 0, 0, 0], [1, 0, 0, 1, 0], [1, 1, 1, 1, 1], [0, 1, 1, 1, 1]]).bool()
    with torch.inference_mode():
        o, a = att(q, kv, kv, inv)
    oo, oa = om.cross_modal_attention(q, kv, kv, inv, asd, "att.", 4)
    for k_, v_ in asd.items():
        out["xattn." + k_] = v_.numpy()
    out["xattn.q"], out["xattn.kv"], out["xattn.invalid"] = q.numpy(), kv.numpy(), inv.numpy()
    rec("xattn.out", o, oo)
    rec("xattn.attn", a, oa)

    # SequenceEncoder(encoder_type='gru') (src/encoders.py:65-77, :137-155), with and without lengths
    gru = ref_encoders.SequenceEncoder(input_dim=104, hidden_dim=256, output_dim=128, num_layers=2, encoder_type="gru",
                                       dropout=0.1)
    gru.load_state_dict({k[len("gru_enc."):]: v for k, v in sd.items() if k.startswith("gru_enc.")}, strict=True)
    gru.eval()
    gru_lengths = torch.tensor([301, 1, 17, 300, 128, 64])
    with torch.inference_mode():
        rec("gru.z", gru(feats), om.sequence_encoder_gru(feats, sd))
        rec("gru.z_lengths", gru(feats, gru_lengths), om.sequence_encoder_gru(feats, sd, lengths=gru_lengths))
    out["gru.lengths"] = gru_lengths.numpy()

    # SequenceEncoder(encoder_type='cnn') (src/encoders.py:79-94, :157-168): eval, and MC passes with the spec masks
    # injected into its one nn.Dropout module, which forward calls twice (sites CNN0, CNN1)
    cnn = ref_encoders.SequenceEncoder(input_dim=104, hidden_dim=256, output_dim=128, num_layers=2, encoder_type="cnn",
                                       dropout=0.1)
    cnn_sd = {k[len("cnn_enc."):]: v for k, v in sd.items() if k.startswith("cnn_enc.")}
    cnn_sd["bn1.num_batches_tracked"] = torch.tensor(0)
    cnn_sd["bn2.num_batches_tracked"] = torch.tensor(0)
    cnn.load_state_dict(cnn_sd, strict=True)
    cnn.eval()
    with torch.inference_mode():
        rec("cnn.z", cnn(feats), om.sequence_encoder_cnn(feats, sd))
    state = {"t": 0, "call": 0}

    def cnn_drop(x, mod=cnn.dropout_layer):
        site = (px.SITE_CNN0, px.SITE_CNN1)[state["call"] % 2]
        state["call"] += 1
        n = int(np.prod(x.shape[1:]))
        keep = px.dropout_keep(n, CLIP_IDS, state["t"], site, mod.p, SEED).reshape(x.shape)
        return x * (torch.from_numpy(keep).float() * float(px.keep_scale(mod.p)))
    cnn.dropout_layer.forward = cnn_drop
    passes = []
    with torch.inference_mode():
        for t in range(3):
            state["t"] = t
            ref_z = cnn(feats)
            passes.append(ref_z)
            ora = om.sequence_encoder_cnn(feats, sd, ctx=om.DropCtx(SEED, t, CLIP_IDS))
            rec(f"cnn.z_pass{t}", ref_z, ora)

    # LateFusionWithUncertainty (src/fusion.py:589-655) behind build_fusion_model('uncertainty'): eval and one stochastic
    # pass with the spec masks injected into its eight nn.Dropout modules
    unc = ref_fusion.build_fusion_model("uncertainty", {"audio": 128, "video": 128}, 8, hidden_dim=256, num_heads=4, dropout=0.3)
    unc.load_state_dict({k[len("unc_model."):]: v for k, v in sd.items() if k.startswith("unc_model.")}, strict=True)
    unc.eval()
    zz = {"audio": z_a, "video": z_v}
    with torch.inference_mode():
        for mname in ("ones", "mixed"):
            lg, info = unc(zz, masks[mname].clone())
            olg, oper, ow = om.late_fusion_with_uncertainty(zz, masks[mname], sd)
            rec(f"unc.logits_{mname}", lg, olg)
            rec(f"unc.per_{mname}", info["per_modality_logits"], oper)
            rec(f"unc.weights_{mname}", info["fusion_weights"], ow)
        lg, info = unc({"video": z_v}, masks["mixed"].clone())                 # audio not given: zeros (:629)
        olg, oper, ow = om.late_fusion_with_uncertainty({"video": z_v}, masks["mixed"], sd)
        rec("unc.logits_video_only", lg, olg)
    sites = {}
    for i, m in enumerate(MODALITIES):
        for j, head in enumerate(("classifiers", "uncert_heads")):
            sites[f"{head}.{m}.0"] = px.SITE_UNC + 4 * i + 2 * j
            sites[f"{head}.{m}.3"] = px.SITE_UNC + 4 * i + 2 * j + 1
    T_UNC = 2
    for name, mod in unc.named_modules():
        if isinstance(mod, nn.Dropout):
            def fwd(x, mod=mod, site=sites[name]):
                n = int(np.prod(x.shape[1:]))
                keep = px.dropout_keep(n, CLIP_IDS, T_UNC, site, mod.p, SEED).reshape(x.shape)
                return x * (torch.from_numpy(keep).float() * float(px.keep_scale(mod.p)))
            mod.forward = fwd
    with torch.inference_mode():
        lg, info = unc(zz, masks["mixed"].clone())
        olg, oper, ow = om.late_fusion_with_uncertainty(zz, masks["mixed"], sd, ctx=om.DropCtx(SEED, T_UNC, CLIP_IDS))
        rec("unc.logits_drop", lg, olg)
        rec("unc.weights_drop", info["fusion_weights"], ow)

    # calibration metrics (src/uncertainty.py:88-183, :512-544) on seeded random logits / labels
    g = torch.Generator().manual_seed(5)
    cal_logits = torch.randn(4000, 8, generator=g) * 2.0
    cal_labels = torch.randint(0, 8, (4000,), generator=g)
    cal_labels[:1500] = cal_logits[:1500].argmax(1)                      # make a third of them correct
    probs = torch.softmax(cal_logits, dim=1)
    confs, preds = probs.max(dim=1)
    ref_cal = {
        "ece": ref_uncertainty.CalibrationMetrics.expected_calibration_error(confs, preds, cal_labels, 15),
        "mce": ref_uncertainty.CalibrationMetrics.maximum_calibration_error(confs, preds, cal_labels, 15),
        "nll": ref_uncertainty.CalibrationMetrics.negative_log_likelihood(cal_logits, cal_labels),
        "accuracy": float((preds == cal_labels).float().mean()),
        "ece10": ref_uncertainty.CalibrationMetrics.expected_calibration_error(confs, preds, cal_labels, 10),
    }
    ora_cal = om.calibration_metrics(cal_logits, cal_labels, 15)
    ora_cal["ece10"] = om.calibration_metrics(cal_logits, cal_labels, 10)["ece"]
    out["cal.logits"], out["cal.labels"] = cal_logits.numpy(), cal_labels.numpy()
    for k_ in ref_cal:
        rec("cal." + k_, torch.tensor(ref_cal[k_], dtype=torch.float64), torch.tensor(ora_cal[k_], dtype=torch.float64))

    print("reference vs oracle/model.py (max abs diff, max |ref|):")
    bad = False
    for k_, (d, s) in worst.items():
        flag = "" if d <= 2e-5 * max(1.0, s) else "   <-- MISMATCH"
        bad |= bool(flag)
        print(f"  {k_:40s} {d:.3e}  {s:.3e}{flag}")
    assert not bad
    path = os.path.join(HERE, "av_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
