"""BASELINE.json `configs` at their full sizes.  The oracle cannot run thousands of clips in seconds, so each config is
checked (a) against the oracle on a seeded SAMPLE of its clips and (b) through size-independent properties over ALL of
them: gain linearity of the front-end, batching / sharding invariance, bit-exact repeatability, the tensor-core path
against the fp32 path, statistics identities of the MC reduction."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import frontend as ofe, model as om, philox as px, synth as osynth

SEED = 42
REL = 1e-4
TC_REL = 2e-3          # stated tolerance of the tensor-core path


def rel_err(got, want):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))


@pytest.fixture(scope="module")
def pkg():
    import multimodal_emotion_detection_b200 as p
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    p._lib.check_device(torch.device("cuda:0"))
    return p


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def sd_np():
    return osynth.synth_weights(SEED)


def _model(pkg, dev, sd_np, precision="fp32"):
    m = pkg.AVHotPath("hybrid")
    m.load_state_dict({k: torch.from_numpy(v) for k, v in osynth.state_dict_for("hybrid", sd_np).items()}, strict=True)
    return m.to(dev).eval().set_precision(precision)


def test_config1_frontend_4096_clips(pkg, dev):
    """configs[1]: log-mel front-end only (n_fft 512, hop 160, 64 mels + 40 MFCC) on 4096 synthetic 3 s clips."""
    B, clip0 = 4096, 20_000
    wave = pkg.synth.synth_audio(B, clip0, SEED, dev)
    feats = pkg.logmel_mfcc(wave)
    assert feats.shape == (B, 301, 104) and torch.isfinite(feats).all()
    # (a) a sample of the clips against the fp64 oracle
    pick = np.array([0, 1, 777, 2048, 4095])
    want = ofe.logmel_mfcc(osynth.synth_audio(pick.astype(np.uint64) + np.uint64(clip0), SEED))
    assert rel_err(feats[pick].cpu().numpy(), want) <= REL
    # (b) batching invariance, bit for bit: the same clips in a batch of 5 (one tile per CTA) and in the batch of 4096
    #     (one clip per CTA walking its tiles)
    small = pkg.logmel_mfcc(wave[pick].contiguous())
    assert torch.equal(small, feats[pick])
    assert torch.equal(pkg.logmel_mfcc(wave), feats)
    # (c) gain linearity over every clip: doubling the waveform (exact in fp32) adds 10 log10(4) dB to every log-mel
    #     value above the log floor and shifts only MFCC 0, by that amount times sqrt(64)
    f2 = pkg.logmel_mfcc(wave * 2.0)
    step = 10.0 * np.log10(4.0)
    lm, lm2 = feats[..., :64], f2[..., :64]
    live = (lm > -90.0).all(dim=-1)                                   # frames away from the 1e-10 floor (zero tails)
    assert float(live.float().mean()) > 0.5
    assert float(((lm2 - lm)[live] - step).abs().max()) <= 2e-3       # dB; fp32 log of a TF32-weighted sum
    mf, mf2 = feats[..., 64:], f2[..., 64:]
    assert float(((mf2 - mf)[live][:, 0] - step * 8.0).abs().max()) <= 1.6e-2
    assert float((mf2 - mf)[live][:, 1:].abs().max()) <= 1e-2
    # silent frames sit on the floor in both
    dead = (lm2 <= -99.999).all(dim=-1)                               # still on the floor after doubling
    assert int(dead.sum()) > 0
    assert torch.equal(f2[dead], feats[dead])


def test_config2_full_forward_batch_512_tensor_core(pkg, dev, sd_np):
    """configs[2]: full audio + video cross-modal attention fusion forward, batch 512, on the tensor-core path."""
    B, clip0 = 512, 40_000
    feats = {"audio": pkg.synth.synth_audio(B, clip0, SEED, dev), "video": pkg.synth.synth_video(B, clip0, SEED, dev)}
    m = _model(pkg, dev, sd_np, "tc")
    lg_tc = m(feats)
    assert lg_tc.shape == (B, 8) and torch.isfinite(lg_tc).all()
    assert torch.equal(m(feats), lg_tc)                               # repeatable bit for bit
    m.set_precision("fp32")
    lg = m(feats)
    assert rel_err(lg_tc.cpu().numpy(), lg.cpu().numpy()) <= TC_REL
    # labels may differ only where the top two fp32 logits are closer than the stated tolerance allows
    diff = (lg_tc.argmax(1) != lg.argmax(1)).nonzero().flatten()
    top2 = lg.topk(2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1])[diff]
    assert len(diff) <= 2 and bool((margin <= 2 * TC_REL * lg.abs().max()).all()), (len(diff), margin)
    # a sample against the fp32 oracle (the reference's arithmetic, pinned by the golden vectors)
    pick = np.array([0, 255, 511])
    ids = pick.astype(np.uint64) + np.uint64(clip0)
    sd = om.to_torch(sd_np)
    want = om.av_forward(torch.from_numpy(ofe.logmel_mfcc(osynth.synth_audio(ids, SEED))),
                         torch.from_numpy(osynth.synth_video(ids, SEED)), None, sd, "hybrid")
    assert rel_err(lg[pick].cpu().numpy(), want.numpy()) <= REL
    assert rel_err(lg_tc[pick].cpu().numpy(), want.numpy()) <= TC_REL
    # batching invariance of the tensor-core path (partial 128-clip tiles): 512 = 200 + 312
    m.set_precision("tc")
    parts = torch.cat([m({k: v[:200] for k, v in feats.items()}), m({k: v[200:] for k, v in feats.items()})])
    assert torch.equal(parts, lg_tc)


def test_config3_uncertainty_T32_batch_256(pkg, dev, sd_np):
    """configs[3]: T = 32 MC-dropout passes, modality dropout p = 0.05, batch 256."""
    B, T, clip0 = 256, 32, 60_000
    feats = {"audio": pkg.synth.synth_audio(B, clip0, SEED, dev), "video": pkg.synth.synth_video(B, clip0, SEED, dev)}
    mask = pkg.synth.modality_mask(B, clip0, 0.05, SEED, dev)
    ids = np.arange(B, dtype=np.uint64) + np.uint64(clip0)
    assert np.array_equal(mask.cpu().numpy(), px.modality_mask(ids, 2, 0.05, SEED))          # bit-exact masks
    assert 0 < int((mask == 0).sum()) < B // 4 and bool((mask.sum(1) >= 1).all())
    m = _model(pkg, dev, sd_np, "fp32")
    mc = pkg.MCDropoutUncertainty(m, num_samples=T)
    mc.seed, mc.clip0 = SEED, clip0
    mean, unc = mc(feats, mask)
    assert mean.shape == (B, 8) and unc.shape == (B,) and bool((unc > 0).all()) and not m.training
    # sharding invariance, bit for bit: two ranks of 128 clips
    pkg.MCDropoutUncertainty._set_dropout_train(m)
    halves = [m.mc_forward({k: v[lo:lo + 128] for k, v in feats.items()}, mask[lo:lo + 128], num_samples=T, seed=SEED,
                           clip0=clip0 + lo) for lo in (0, 128)]
    assert torch.equal(torch.cat([h[0] for h in halves]), mean) and torch.equal(torch.cat([h[1] for h in halves]), unc)
    packed = torch.cat([h[3] for h in halves])
    assert torch.equal(packed[:, :8], mean) and torch.equal(packed[:, 8], unc)
    assert torch.equal(torch.cat([h[2] for h in halves]).long(), mean.argmax(1))
    # a sample against the oracle running the reference's procedure (T full forwards, spec masks)
    pick = np.array([0, 100, 255])
    sd = om.to_torch(sd_np)
    o_mean, o_unc, _ = om.mc_dropout_forward(torch.from_numpy(ofe.logmel_mfcc(osynth.synth_audio(ids[pick], SEED))),
                                             torch.from_numpy(osynth.synth_video(ids[pick], SEED)),
                                             torch.from_numpy(px.modality_mask(ids[pick], 2, 0.05, SEED)), sd, T, SEED,
                                             ids[pick], "hybrid")
    assert rel_err(mean[pick].cpu().numpy(), o_mean.numpy()) <= REL
    assert rel_err(unc[pick].cpu().numpy(), o_unc.numpy()) <= 1e-3
    assert np.array_equal(mean[pick].argmax(1).cpu().numpy(), o_mean.argmax(1).numpy())
    # the tensor-core path on the same batch: same masks, stated tolerance
    m.set_precision("tc")
    mean_tc, unc_tc, _, _ = m.mc_forward(feats, mask, num_samples=T, seed=SEED, clip0=clip0)
    assert rel_err(mean_tc.cpu().numpy(), mean.cpu().numpy()) <= TC_REL
    assert rel_err(unc_tc.cpu().numpy(), unc.cpu().numpy()) <= 5e-2


def test_config4_sweep_shards_agree_with_the_whole(pkg, dev, sd_np):
    """configs[4]: the 1 M-clip sweep is index-sharded; a window of it computed as 1, 2, 4 and 8 shards
    (parallel.shard_range) gives the same packed records bit for bit, at clip indices near the end of the sweep."""
    from multimodal_emotion_detection_b200 import parallel as par
    total, window, T = 1_000_000, 192, 4
    m = _model(pkg, dev, sd_np, "tc")
    pkg.MCDropoutUncertainty._set_dropout_train(m)

    def run(lo, hi):
        n = hi - lo
        f = {"audio": pkg.synth.synth_audio(n, lo, SEED, dev), "video": pkg.synth.synth_video(n, lo, SEED, dev)}
        return m.mc_forward(f, pkg.synth.modality_mask(n, lo, 0.05, SEED, dev), num_samples=T, seed=SEED, clip0=lo)[3]

    base = total - window
    whole = run(base, total)
    for world in (2, 4, 8):
        ranges = [par.shard_range(window, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == window and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        got = torch.cat([run(base + lo, base + hi) for lo, hi in ranges])
        assert torch.equal(got, whole), world
    for world in (1, 2, 4, 8):                                         # the full sweep partitions exactly
        r = [par.shard_range(total, k, world) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == total and all(a[1] == b[0] for a, b in zip(r, r[1:]))
