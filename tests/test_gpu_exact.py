"""Label exactness of the tensor-core path (exact.NearTieResolver): the near-tie bookkeeping kernels against numpy, and
labels of precision 'tc' == labels of the fp32 CUDA path on every batch for several weight seeds (north star: predicted
labels identical to the reference, whose prediction is argmax of the fp32 mean logits, src/uncertainty.py:77)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from test_gpu_parity import SEED, pkg, dev  # noqa: F401  (fixtures)

P_MOD = 0.05


def test_flag_gather_scatter_vs_numpy(pkg, dev):
    lib, L = pkg._lib.load(), pkg._lib
    rng = np.random.default_rng(3)
    B, Cn, cap = 1000, 8, 128
    mean = rng.normal(size=(B, Cn)).astype(np.float32)
    mean[17] = mean[17, 0]                                   # exact tie
    mean[400, 3] = np.nan                                    # NaN rows are flagged
    margin_abs = 0.04
    srt = np.sort(np.where(np.isnan(mean), -np.inf, mean), axis=1)
    margin = srt[:, -1] - srt[:, -2]
    want = set(np.nonzero(~(margin >= np.float32(margin_abs)) | np.isnan(mean).any(1))[0].tolist())
    assert 2 <= len(want) <= cap // 2
    d_mean = torch.from_numpy(mean).to(dev)
    state = torch.zeros(4, dtype=torch.int32, device=dev)
    ent = torch.zeros(3, cap, dtype=torch.int32, device=dev)
    rows = torch.from_numpy(rng.integers(0, 255, size=(B, 48), dtype=np.uint8)).to(dev)
    pool = torch.zeros(cap, 48, dtype=torch.uint8, device=dev)
    s = L.stream_ptr(dev)
    for call, clip0 in ((5, 1000), (6, 5000)):              # two calls of one window append one after the other
        L.check(lib.avhot_flag_near_ties(L.ptr(d_mean), B, Cn, margin_abs, call, clip0, L.ptr(state), L.ptr(ent[0]),
                                         L.ptr(ent[1]), L.ptr(ent[2]), cap, s))
        L.check(lib.avhot_gather_flagged(L.ptr(rows), 48, B, L.ptr(state), L.ptr(ent[1]), L.ptr(pool), cap, s))
    st = state.cpu().numpy()
    n = int(st[0])
    assert n == 2 * len(want) and st[1] == 0
    e = ent.cpu().numpy()[:, :n]
    for k, (call, clip0) in enumerate(((5, 1000), (6, 5000))):
        sel = e[0] == call
        assert sel.sum() == len(want) and set(e[1, sel].tolist()) == want
        assert (e[2, sel] == e[1, sel] + clip0).all()
        assert (np.nonzero(sel)[0] // len(want) == k).all()          # the second call's entries come after the first's
    assert (pool.cpu().numpy()[:n] == rows.cpu().numpy()[e[1]]).all()
    # scatter: records of call 6 only
    r_packed = torch.from_numpy(rng.normal(size=(n, Cn + 1)).astype(np.float32)).to(dev)
    r_labels = torch.arange(n, dtype=torch.int32, device=dev)
    o_mean = torch.zeros(B, Cn, device=dev); o_unc = torch.zeros(B, device=dev)
    o_lab = torch.full((B,), -1, dtype=torch.int32, device=dev); o_packed = torch.zeros(B, Cn + 1, device=dev)
    L.check(lib.avhot_scatter_resolved(L.ptr(ent[0]), L.ptr(ent[1]), n, 6, L.ptr(r_packed), L.ptr(r_labels), L.ptr(o_mean),
                                       L.ptr(o_unc), L.ptr(o_lab), L.ptr(o_packed), Cn, s))
    lab = o_lab.cpu().numpy()
    assert set(np.nonzero(lab >= 0)[0].tolist()) == want
    for idx in np.nonzero(e[0] == 6)[0]:
        row = e[1, idx]
        assert lab[row] == idx
        assert (o_packed[row].cpu().numpy() == r_packed[idx].cpu().numpy()).all()
        assert (o_mean[row].cpu().numpy() == r_packed[idx, :Cn].cpu().numpy()).all()
        assert float(o_unc[row]) == float(r_packed[idx, Cn])
    # overflow: capacity smaller than the flagged set raises the flag and never writes past the arrays
    state.zero_()
    L.check(lib.avhot_flag_near_ties(L.ptr(d_mean), B, Cn, 10.0, 1, 0, L.ptr(state), L.ptr(ent[0]), L.ptr(ent[1]),
                                     L.ptr(ent[2]), cap, s))
    st = state.cpu().numpy()
    assert st[0] == B and st[1] == 1


def _batch(pkg, dev, B, clip0):
    feats = {"audio": pkg.synth.synth_audio(B, clip0, SEED, dev), "video": pkg.synth.synth_video(B, clip0, SEED, dev)}
    return feats, pkg.synth.modality_mask(B, clip0, P_MOD, SEED, dev)


@pytest.mark.parametrize("weight_seed", [42, 7, 123])
def test_tc_labels_equal_fp32_labels_every_batch(pkg, dev, weight_seed):
    """Four batches of 768 clips per weight seed through the tensor-core path with the resolver in deferred mode (as
    bench.py uses it) and in immediate mode (mc_forward's default): labels identical to the fp32 CUDA path, and the
    patched records are the fp32 path's bit for bit."""
    torch.manual_seed(weight_seed)
    model = pkg.AVHotPath("hybrid").to(dev).eval().set_precision("tc")
    pkg.MCDropoutUncertainty._set_dropout_train(model)
    B, T = 768, 8
    res = pkg.exact.NearTieResolver(model, capacity=512)
    outs, batches = [], []
    for j in range(4):
        clip0 = 10_000 + j * B
        feats, mask = _batch(pkg, dev, B, clip0)
        raw = model.mc_forward(feats, mask, num_samples=T, seed=SEED, clip0=clip0, exact_labels=False)
        out = model.mc_forward(feats, mask, num_samples=T, seed=SEED, clip0=clip0, exact_labels=res)
        outs.append((raw, out))
        batches.append((feats, mask, clip0))
        if j == 1:
            res.flush()                                                   # windows of two batches
    n_last = res.flush()
    assert res.pending_calls == 0 and res.flagged_total >= n_last
    assert res.flagged_total <= 0.05 * res.clips_total, "near ties should be a small fraction of the clips"
    model.set_precision("fp32")
    raw_flips = 0
    for (raw, out), (feats, mask, clip0) in zip(outs, batches):
        ref = model.mc_forward(feats, mask, num_samples=T, seed=SEED, clip0=clip0)
        raw_flips += int((raw[2] != ref[2]).sum())
        assert int((out[2] != ref[2]).sum()) == 0, "label differs from the fp32 path after near-tie resolution"
        changed = (out[3] != raw[3]).any(dim=1)                           # rows the resolver patched
        assert torch.equal(out[3][changed], ref[3][changed]) and torch.equal(out[0][changed], ref[0][changed])
        assert torch.equal(out[1][changed], ref[1][changed])
        assert torch.equal(out[0].argmax(1).int(), out[2])
        gmax = float(ref[0].abs().max())
        assert float((out[0] - ref[0]).abs().max()) <= 2e-3 * gmax        # the stated tolerance holds on every clip
        top = ref[0].topk(2, dim=1)                                        # the resolver's premise: the error of the top-2
        d = raw[0] - ref[0]                                                # difference stays inside the margin
        dd = (d.gather(1, top.indices[:, :1]) - d.gather(1, top.indices[:, 1:2])).abs().max()
        assert float(dd) < 2.0 * res.margin        # sanity (the measured envelope is ~0.9 of the margin at 8 sigma)
    print(f"weight seed {weight_seed}: {res.flagged_total} of {res.clips_total} clips re-evaluated; the unresolved "
          f"tensor-core path had {raw_flips} label flips")
    # immediate mode (the default of mc_forward with precision 'tc')
    model.set_precision("tc")
    feats, mask, clip0 = batches[0]
    imm = model.mc_forward(feats, mask, num_samples=T, seed=SEED, clip0=clip0)
    assert torch.equal(imm[3], outs[0][1][3]) and torch.equal(imm[2], outs[0][1][2])


def test_resolver_overflow_falls_back_to_fp32(pkg, dev):
    """More near ties than the pool holds (forced here with a margin as wide as the logit scale): the immediate mode
    evaluates the whole batch in fp32 instead of failing - bit-identical to the fp32 path -, the deferred mode reports
    the overflow.  A classifier with all-zero weights has logit scale 0: nothing is ambiguous, nothing is flagged."""
    torch.manual_seed(1)
    model = pkg.AVHotPath("hybrid").to(dev).eval().set_precision("tc")
    pkg.MCDropoutUncertainty._set_dropout_train(model)
    B, T = 300, 4
    feats, mask = _batch(pkg, dev, B, 0)
    model.near_tie_margin_rel = 1.0
    out = model.mc_forward(feats, mask, num_samples=T, seed=SEED, clip0=0)
    assert model.precision == "tc"
    model.set_precision("fp32")
    ref = model.mc_forward(feats, mask, num_samples=T, seed=SEED, clip0=0)
    model.set_precision("tc")
    assert torch.equal(out[3], ref[3]) and torch.equal(out[2], ref[2])
    res = pkg.exact.NearTieResolver(model, capacity=16, margin_rel=1.0)
    model.mc_forward(feats, mask, num_samples=T, seed=SEED, clip0=0, exact_labels=res)
    with pytest.raises(RuntimeError):
        res.flush()
    with torch.no_grad():
        model.fusion_model.classifier.weight.zero_()
        model.fusion_model.classifier.bias.fill_(1.0)
    res0 = pkg.exact.NearTieResolver(model, capacity=16)
    assert res0.margin == 0.0
    out = model.mc_forward(feats, mask, num_samples=T, seed=SEED, clip0=0, exact_labels=res0)
    assert res0.flush() == 0 and int(out[2].abs().sum()) == 0 and bool((out[0] == 1.0).all())
