#!/usr/bin/env python
"""Which stage of the tensor-core path the error of the top-2 mean-logit difference comes from: the same batch with one
stage at a time switched to the fp32 kernels (run on a B200).  python tools/measure_error_sources.py > gpurun_out/err_sources.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import multimodal_emotion_detection_b200 as pkg  # noqa: E402
from multimodal_emotion_detection_b200.exact import logit_scale  # noqa: E402

dev = torch.device("cuda:0")
SEED, B, T = 42, 148 * 128, 32
torch.manual_seed(42)
model = pkg.AVHotPath("hybrid").to(dev).eval()
pkg.MCDropoutUncertainty._set_dropout_train(model)
S = logit_scale(model)
feats = {"audio": pkg.synth.synth_audio(B, 0, SEED, dev), "video": pkg.synth.synth_video(B, 0, SEED, dev)}
mask = pkg.synth.modality_mask(B, 0, 0.05, SEED, dev)
model.set_precision("fp32")
ref = model.mc_forward(feats, mask, num_samples=T, seed=SEED, clip0=0)[0]
top = ref.topk(2, dim=1)
TC, F32 = pkg._lib.PREC_TC, pkg._lib.PREC_FP32
out = {"logit_scale": S, "clips": B}
for name, (pa, pv, pf) in {"all_tc": (TC, TC, TC), "audio_fp32": (F32, TC, TC), "video_fp32": (TC, F32, TC),
                           "fusion_fp32": (TC, TC, F32), "only_audio_tc": (TC, F32, F32), "only_video_tc": (F32, TC, F32),
                           "only_fusion_tc": (F32, F32, TC)}.items():
    model.set_precision("tc")
    model.encoders["audio"].precision, model.encoders["video"].precision, model.fusion_model.precision = pa, pv, pf
    model.fold_video_projection = pf == TC and pv == TC
    a = model.mc_forward(feats, mask, num_samples=T, seed=SEED, clip0=0, exact_labels=False)[0]
    d = a - ref
    dd = (d.gather(1, top.indices[:, :1]) - d.gather(1, top.indices[:, 1:2])).squeeze(1) / S
    out[name] = {"top2_diff_err_rms": float(dd.pow(2).mean().sqrt()), "top2_diff_err_max": float(dd.abs().max()),
                 "q99.9": float(dd.abs().quantile(0.999)), "logit_err_max_rel_gmax": float(d.abs().max() / ref.abs().max())}
print(json.dumps(out, indent=1))
