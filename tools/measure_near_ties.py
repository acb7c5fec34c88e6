#!/usr/bin/env python
"""Measurement behind exact.DEFAULT_MARGIN_REL (run on a B200): distribution of the top-2 margins of the mean logits,
of the tensor-core path's error on the top-2 difference, and what a flush of the fp32 kernels costs at several sizes.
    python tools/measure_near_ties.py [--batches 3] > gpurun_out/near_ties.json"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import multimodal_emotion_detection_b200 as pkg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batches", type=int, default=3)
ap.add_argument("--batch", type=int, default=148 * 128)
ap.add_argument("--passes", type=int, default=32)
ap.add_argument("--weight-seed", type=int, default=42)
args = ap.parse_args()
dev = torch.device("cuda:0")
SEED = 42
torch.manual_seed(args.weight_seed)
model = pkg.AVHotPath("hybrid").to(dev).eval()
pkg.MCDropoutUncertainty._set_dropout_train(model)
out = {"weight_seed": args.weight_seed, "batches": []}
B, T = args.batch, args.passes
for j in range(args.batches):
    c0 = j * B
    feats = {"audio": pkg.synth.synth_audio(B, c0, SEED, dev), "video": pkg.synth.synth_video(B, c0, SEED, dev)}
    mask = pkg.synth.modality_mask(B, c0, 0.05, SEED, dev)
    model.set_precision("tc")
    a = model.mc_forward(feats, mask, num_samples=T, seed=SEED, clip0=c0, exact_labels=False)
    model.set_precision("fp32")
    b = model.mc_forward(feats, mask, num_samples=T, seed=SEED, clip0=c0)
    gmax = float(b[0].abs().max())
    top = b[0].topk(2, dim=1)
    margin = (top.values[:, 0] - top.values[:, 1]) / gmax
    d = (a[0] - b[0])
    dd = (d.gather(1, top.indices[:, :1]) - d.gather(1, top.indices[:, 1:2])).squeeze(1) / gmax   # error of the top-2 difference
    flips = (a[2] != b[2])
    q = torch.tensor([0.5, 0.9, 0.99, 0.999, 0.9999], device=dev)
    ent = {"clips": B, "gmax": gmax, "flips_unresolved": int(flips.sum()),
           "err_max_rel": float(d.abs().max() / gmax), "err_rms_rel": float(d.pow(2).mean().sqrt() / gmax),
           "top2_diff_err_rms": float(dd.pow(2).mean().sqrt()), "top2_diff_err_max": float(dd.abs().max()),
           "top2_diff_err_quantiles": dict(zip(["50", "90", "99", "99.9", "99.99"], [float(x) for x in dd.abs().quantile(q)])),
           "margin_quantiles": dict(zip(["50", "90", "99", "99.9", "99.99"], [float(x) for x in margin.quantile(1 - q)])),
           "frac_margin_below": {str(m): float((margin < m).float().mean()) for m in (2e-4, 4e-4, 6e-4, 8e-4, 1e-3, 2e-3, 4e-3)},
           "max_margin_of_a_flip": float(margin[flips].max()) if bool(flips.any()) else None}
    out["batches"].append(ent)
    del feats, mask, a, b
# cost of the fp32 path at flush sizes
model.set_precision("fp32")
cost = {}
for n in (256, 1024, 2560, 5120):
    feats = {"audio": pkg.synth.synth_audio(n, 0, SEED, dev), "video": pkg.synth.synth_video(n, 0, SEED, dev)}
    mask = pkg.synth.modality_mask(n, 0, 0.05, SEED, dev)
    ids = torch.arange(n, device=dev, dtype=torch.int32)
    model._run(feats, mask, T, True, SEED, 0, 0, clip_ids=ids)
    pkg._lib.prof_reset(); pkg._lib.prof_enable(True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    model._run(feats, mask, T, True, SEED, 0, 0, clip_ids=ids)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    pkg._lib.prof_enable(False)
    cost[str(n)] = {"ms": dt * 1e3, "us_per_clip": dt * 1e6 / n, "stages_ms": {k: v[0] for k, v in pkg._lib.prof_report().items()}}
out["fp32_flush_cost"] = cost
print(json.dumps(out, indent=1))
