"""Parity tests proper: the CUDA path (through the C ABI) against the oracle and the reference-generated golden
vectors, on the same seeded inputs.  Tolerances (BASELINE.json north_star): features and logits within 1e-4
relative in fp32; dropout masks, synthetic inputs and predicted labels bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import frontend as ofe, model as om, philox as px, synth as osynth

SEED = 42
REL = 1e-4


def rel_err(got, want):
    """norm-wise relative error max|got - want| / max|want| (the '1e-4 relative' of the north star)."""
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
def world(golden):
    """Seeded inputs / weights shared by the tests (oracle side, CPU)."""
    ids = golden["clip_ids"]
    sd_np = osynth.synth_weights(SEED)
    wave = osynth.synth_audio(ids, SEED)
    frames_u8 = osynth.synth_video_u8(ids, SEED)
    frames = frames_u8.astype(np.float32) / np.float32(255.0)
    feats = ofe.logmel_mfcc(wave)
    return dict(ids=ids, sd_np=sd_np, sd=om.to_torch(sd_np), sd64=om.to_torch(sd_np, torch.float64), wave=wave,
                frames=frames, frames_u8=frames_u8, feats=feats)


def make_model(pkg, dev, world, head="hybrid"):
    m = pkg.AVHotPath(head)
    m.load_state_dict(osynth.state_dict_for(head, world["sd"]), strict=True)
    return m.to(dev).eval()


# ---------------------------------------------------------------------------------------------------------
# bit-exact pieces: synthetic inputs, masks
# ---------------------------------------------------------------------------------------------------------
def test_synth_inputs_bit_exact(pkg, dev):
    for clip0, B in ((0, 5), (1000003, 3), (4294967294, 2)):
        ids = (np.arange(B, dtype=np.uint64) + np.uint64(clip0))
        a = pkg.synth.synth_audio(B, clip0, SEED, dev).cpu().numpy()
        assert (a == osynth.synth_audio(ids, SEED)).all()
        v8 = pkg.synth.synth_video(B, clip0, SEED, dev, dtype=torch.uint8).cpu().numpy()
        assert (v8 == osynth.synth_video_u8(ids, SEED)).all()
        v = pkg.synth.synth_video(B, clip0, SEED, dev).cpu().numpy()
        assert (v == osynth.synth_video(ids, SEED)).all()
    a = pkg.synth.synth_audio(4, 7, SEED, dev)
    assert float(a.abs().max(dim=1).values.min()) == 1.0          # peak-normalised (src/dataprocessing.py:207-208)


def test_dropout_masks_bit_exact(pkg, dev):
    ids = np.array([0, 1, 2, 77], dtype=np.uint64)
    for site, n, p in ((px.SITE_FRAME_MLP, 16 * 256, 0.1), (px.SITE_FRAME_POOL, 256, 0.1), (px.SITE_GATE, 256, 0.3),
                       (px.SITE_ATTN, 8, 0.3), (px.SITE_ATTN + 1, 8, 0.3), (px.SITE_GATE, 7, 0.5)):
        for t in (0, 5, 31):
            got = pkg.synth.dropout_keep(3, n, 0, t, site, p, SEED, dev).cpu().numpy().astype(bool)
            assert (got == px.dropout_keep(n, ids[:3], t, site, p, SEED)).all()
    got = pkg.synth.dropout_keep(1, 256, 77, 2, px.SITE_GATE, 0.3, 123456789012345, dev).cpu().numpy().astype(bool)
    assert (got == px.dropout_keep(256, ids[3:], 2, px.SITE_GATE, 0.3, 123456789012345)).all()


def test_modality_mask_bit_exact(pkg, dev):
    ids = np.arange(5000, dtype=np.uint64) + np.uint64(123)
    for p in (0.0, 0.05, 0.9):
        got = pkg.synth.modality_mask(5000, 123, p, SEED, dev).cpu().numpy()
        assert (got == px.modality_mask(ids, 2, p, SEED)).all()


def test_condition_audio(pkg, dev):
    rng = np.random.default_rng(0)
    x = rng.normal(size=(4, 5000)).astype(np.float32)
    x[2] = 0.0                                                     # silent clip: no division (:207)
    lengths = np.array([5000, 1234, 5000, 1], dtype=np.int32)
    for n_out in (3000, 5000, 8000):
        got = pkg.condition_audio(torch.from_numpy(x).to(dev), torch.from_numpy(lengths).to(dev), n_out).cpu().numpy()
        for b in range(4):
            L = min(int(lengths[b]), n_out)
            y = np.zeros(n_out, dtype=np.float32)
            y[:L] = x[b, :L]
            peak = np.abs(y).max()
            if peak > 0:
                y = y / peak
            assert (got[b] == y).all()


def test_condition_audio_pcm16_equals_float_path(pkg, dev):
    """16-bit PCM in (what the .wav holds) == converting to k / 32768 first (librosa.load, src/dataprocessing.py:196) and
    conditioning the floats: bit-identical, including the -32768 sample, a silent clip and truncation / padding."""
    rng = np.random.default_rng(1)
    k = rng.integers(-20000, 20000, size=(5, 6000)).astype(np.int16)
    k[0, 17] = -32768
    k[2] = 0
    k[3] //= 300                                                   # quiet clip: the peak is far from full scale
    lengths = np.array([6000, 4321, 6000, 6000, 1], dtype=np.int32)
    for n_out in (4000, 6000, 7001):
        a = pkg.condition_audio(torch.from_numpy(k).to(dev), torch.from_numpy(lengths).to(dev), n_out)
        b = pkg.condition_audio(torch.from_numpy(k.astype(np.float32) / 32768.0).to(dev), torch.from_numpy(lengths).to(dev), n_out)
        assert a.dtype == torch.float32 and torch.equal(a, b)
        assert float(a[2].abs().max()) == 0.0 and float(a[0].abs().max()) == 1.0


# ---------------------------------------------------------------------------------------------------------
# a2 front-end
# ---------------------------------------------------------------------------------------------------------
def test_frontend_vs_oracle(pkg, dev, world, golden):
    got = pkg.logmel_mfcc(torch.from_numpy(world["wave"]).to(dev)).cpu().numpy()
    want = world["feats"]
    assert got.shape == want.shape == (6, 301, 104)
    assert rel_err(got, want) <= REL
    assert rel_err(got[..., :64], want[..., :64]) <= REL and rel_err(got[..., 64:], want[..., 64:]) <= REL
    assert np.abs(got - want).max() <= 5e-3                        # dB / DCT-of-dB units
    assert rel_err(got[0], golden["feats_clip0"]) <= REL and rel_err(got[-1], golden["feats_last"]) <= REL
    mel, dct, win = pkg.dataprocessing.frontend_tables()
    np.testing.assert_allclose(mel, ofe.mel_filterbank(), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(dct, ofe.dct_matrix(), rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(win, ofe.hann_periodic(), rtol=0, atol=1e-7)


@pytest.mark.parametrize("n", [257, 1000, 1600, 16000, 48000, 48160])
def test_frontend_ragged_lengths_and_silence(pkg, dev, n):
    rng = np.random.default_rng(n)
    w = rng.uniform(-1, 1, size=(3, n)).astype(np.float32)
    w[1] = 0.0                                                     # silence -> log floor
    w[2, n // 2:] = 0.0                                            # zero tail (right padding)
    got = pkg.logmel_mfcc(torch.from_numpy(w).to(dev)).cpu().numpy()
    want = ofe.logmel_mfcc(w)
    assert got.shape == want.shape == (3, 1 + n // 160, 104)
    assert rel_err(got, want) <= REL
    assert np.allclose(got[1, :, :64], -100.0, atol=1e-4)


def test_frontend_rejects_bad_input(pkg, dev):
    with pytest.raises(ValueError):
        pkg.logmel_mfcc(torch.zeros(2, 100, device=dev))           # shorter than the reflect pad
    with pytest.raises(ValueError):
        pkg.logmel_mfcc(torch.zeros(2, 3, 4, 5, device=dev))
    assert pkg.logmel_mfcc(torch.zeros(0, 48000, device=dev)).shape == (0, 301, 104)
    assert pkg.logmel_mfcc(torch.zeros(2, 48000, 1, device=dev)).shape == (2, 301, 104)


# ---------------------------------------------------------------------------------------------------------
# a3 LSTM encoder
# ---------------------------------------------------------------------------------------------------------
def test_sequence_encoder_vs_oracle_and_golden(pkg, dev, world, golden):
    m = make_model(pkg, dev, world)
    feats = torch.from_numpy(world["feats"])
    got = m.encoders["audio"](feats.to(dev)).cpu().numpy()
    want64 = om.sequence_encoder_lstm(feats.double(), world["sd64"]).numpy()
    assert rel_err(got, want64) <= REL
    assert rel_err(got, golden["z_audio"]) <= REL                  # the reference's own nn.LSTM output


def test_sequence_encoder_lengths_and_small_shapes(pkg, dev, world):
    m = make_model(pkg, dev, world)
    enc = m.encoders["audio"]
    rng = np.random.default_rng(3)
    x = rng.normal(size=(5, 12, 104)).astype(np.float32)
    lengths = np.array([12, 1, 7, 12, 3])
    got = enc(torch.from_numpy(x).to(dev), torch.from_numpy(lengths)).cpu().numpy()
    for b in range(5):                                             # packed-sequence semantics (:139-147)
        want = om.sequence_encoder_lstm(torch.from_numpy(x[b:b + 1, :lengths[b]]).double(), world["sd64"]).numpy()
        assert rel_err(got[b:b + 1], want) <= REL
    full = enc(torch.from_numpy(x).to(dev)).cpu().numpy()
    want = om.sequence_encoder_lstm(torch.from_numpy(x).double(), world["sd64"]).numpy()
    assert rel_err(full, want) <= REL
    big = enc(torch.from_numpy(np.tile(x, (80, 1, 1))).to(dev)).cpu().numpy()   # B = 400: the 32-clip tile kernel
    assert rel_err(big[:5], want) <= REL and (big[5:10] == big[:5]).all()
    assert enc(torch.zeros(0, 12, 104, device=dev)).shape == (0, 128)
    with pytest.raises(ValueError):
        enc(torch.zeros(2, 12, 100, device=dev))


# ---------------------------------------------------------------------------------------------------------
# a4 frame encoder
# ---------------------------------------------------------------------------------------------------------
def test_frame_encoder_eval(pkg, dev, world, golden):
    m = make_model(pkg, dev, world)
    enc = m.encoders["video"]
    frames = torch.from_numpy(world["frames"])
    hid = enc.hidden(frames.to(dev))
    want_hid = om.frame_hidden(frames.double(), world["sd64"]).numpy()
    assert rel_err(hid.cpu().numpy(), want_hid) <= REL
    hid8 = enc.hidden(torch.from_numpy(world["frames_u8"]).to(dev))                 # u8 ingest == fp32 k/255
    assert (hid8 == hid).all()
    got = enc(frames.to(dev)).cpu().numpy()
    assert rel_err(got, golden["z_video"]) <= REL
    assert rel_err(got, om.frame_encoder(frames.double(), world["sd64"]).numpy()) <= REL


def test_frame_encoder_dropout_passes_and_mask(pkg, dev, world):
    m = make_model(pkg, dev, world)
    enc = m.encoders["video"]
    frames = torch.from_numpy(world["frames"])
    hid = enc.hidden(frames.to(dev))
    hid64 = om.frame_hidden(frames.double(), world["sd64"])
    T = 5
    z = enc.tail(hid, None, T, 0.1, SEED, 0, 0).cpu().numpy()      # clips 0..5 are ids 0..5 only for the first 4
    ids = np.arange(6, dtype=np.uint64)
    for t in range(T):
        ctx = om.DropCtx(seed=SEED, t=t, clip_ids=ids, p_encoder=0.1)
        want = om.frame_encoder_tail(hid64, world["sd64"], ctx=ctx).numpy()
        assert rel_err(z[:, t], want) <= REL
    z2 = enc.tail(hid[2:4], None, 2, 0.1, SEED, 2, 3).cpu().numpy()   # clip0 / t0 offsets select the same masks
    assert (z2 == z[2:4, 3:5]).all()
    # forward(frames, mask): invalid frames get no weight (src/encoders.py:321-324)
    valid = torch.ones(6, 16, dtype=torch.bool)
    valid[1, 5:] = False
    valid[2, :15] = False
    got = enc(frames.to(dev), valid.to(dev)).cpu().numpy()
    x = hid64
    scores = (x @ world["sd64"]["encoders.video.attention.weight"].T + world["sd64"]["encoders.video.attention.bias"]).squeeze(-1)
    scores = scores.masked_fill(~valid, torch.finfo(torch.float32).min)
    pooled = (torch.softmax(scores, dim=1).unsqueeze(-1) * x).sum(dim=1)
    y = om._layer_norm(pooled, world["sd64"]["encoders.video.projection.0.weight"], world["sd64"]["encoders.video.projection.0.bias"])
    want = om._lin(y, world["sd64"], "encoders.video.projection.1").numpy()
    assert rel_err(got, want) <= REL


@pytest.mark.parametrize("pooling", ["average", "max"])
def test_frame_encoder_average_and_max_pooling(pkg, dev, world, pooling):
    """FrameEncoder(temporal_pooling='average' | 'max'), src/encoders.py:276-293: eval, frame mask, and MC passes."""
    torch.manual_seed(3)
    enc = pkg.FrameEncoder(12544, hidden_dim=256, output_dim=128, temporal_pooling=pooling, dropout=0.1).to(dev).eval()
    assert not hasattr(enc, "attention")
    sd64 = {"v." + k: v.detach().double().cpu() for k, v in enc.state_dict().items()}
    frames = torch.from_numpy(world["frames"])
    hid64 = om.frame_hidden(frames.double(), sd64, "v.")
    got = enc(frames.to(dev)).cpu().numpy()
    assert rel_err(got, om.frame_encoder_tail(hid64, sd64, "v.", pooling=pooling).numpy()) <= REL
    valid = torch.ones(6, 16, dtype=torch.bool)
    valid[1, 5:] = False
    valid[2, :15] = False
    valid[3, :] = False if pooling == "average" else True          # average: all frames invalid -> zeros / clamp_min(1)
    got = enc(frames.to(dev), valid.to(dev)).cpu().numpy()
    assert rel_err(got, om.frame_encoder_tail(hid64, sd64, "v.", pooling=pooling, valid=valid).numpy()) <= REL
    hid = enc.hidden(frames.to(dev))
    T = 3
    z = enc.tail(hid, None, T, 0.1, SEED, 0, 0).cpu().numpy()
    for t in range(T):
        ctx = om.DropCtx(seed=SEED, t=t, clip_ids=np.arange(6, dtype=np.uint64), p_encoder=0.1)
        assert rel_err(z[:, t], om.frame_encoder_tail(hid64, sd64, "v.", ctx=ctx, pooling=pooling).numpy()) <= REL


def test_attention_pool_kernel(pkg, dev, world):
    """FrameEncoder.attention_pool (src/encoders.py:300-327) as a device kernel, with and without a frame mask."""
    m = make_model(pkg, dev, world)
    enc = m.encoders["video"]
    rng = np.random.default_rng(8)
    x = torch.from_numpy(rng.normal(size=(37, 11, 256)).astype(np.float32))
    valid = torch.from_numpy(rng.uniform(size=(37, 11)) > 0.3)
    valid[:, 0] = True
    for v in (None, valid):
        got = enc.attention_pool(x.to(dev), None if v is None else v.to(dev)).cpu().numpy()
        want = om.attention_pool(x.double(), world["sd64"], valid=v).numpy()
        assert got.shape == (37, 256) and rel_err(got, want) <= REL
    with pytest.raises(ValueError):
        enc.attention_pool(torch.zeros(2, 3, 100, device=dev))


# ---------------------------------------------------------------------------------------------------------
# a5 / a6 / a7 fusion
# ---------------------------------------------------------------------------------------------------------
def _masks(world):
    ids = world["ids"]
    return {"ones": np.ones((6, 2), np.float32),
            "mixed": np.array([[1, 1], [1, 0], [0, 1], [0, 0], [1, 1], [0, 1]], np.float32),
            "p05": px.modality_mask(ids, 2, 0.05, SEED)}


def test_hybrid_fusion_vs_reference_golden(pkg, dev, world, golden):
    m = make_model(pkg, dev, world)
    fm = m.fusion_model
    feats = {"audio": torch.from_numpy(golden["z_audio"]).to(dev), "video": torch.from_numpy(golden["z_video"]).to(dev)}
    for name, mask in _masks(world).items():
        logits, info = fm(feats, torch.from_numpy(mask).to(dev), return_attention=True)
        assert rel_err(logits.cpu().numpy(), golden[f"logits_hybrid_{name}"]) <= REL
        assert rel_err(info["fusion_weights"].cpu().numpy(), golden[f"fusion_weights_{name}"]) <= REL
        assert rel_err(info["per_modality_attention"]["audio"].cpu().numpy(), golden[f"attn_audio_{name}"]) <= REL
        assert rel_err(info["per_modality_attention"]["video"].cpu().numpy(), golden[f"attn_video_{name}"]) <= REL
        assert info["H_att"].shape == (6, 2, 256)
        assert info["per_modality_attention"]["audio"].shape == (6, 4, 1, 2)
        only = fm(feats, torch.from_numpy(mask).to(dev))          # no diagnostics: the hoisted kernels
        assert isinstance(only, torch.Tensor)
        assert rel_err(only.cpu().numpy(), golden[f"logits_hybrid_{name}"]) <= REL
        assert rel_err(only.cpu().numpy(), logits.cpu().numpy()) <= 1e-5
    none_mask = fm(feats).cpu().numpy()                            # modality_mask=None == ones (:468-469)
    assert rel_err(none_mask, golden["logits_hybrid_ones"]) <= REL
    # absent modality: zero features + masked (:480-485)
    got = fm({"audio": feats["audio"], "video": None}).cpu().numpy()
    z = {"audio": torch.from_numpy(golden["z_audio"]).double(), "video": torch.zeros(6, 128, dtype=torch.float64)}
    want = om.hybrid_fusion(z, torch.tensor([[1.0, 0.0]] * 6), world["sd64"]).numpy()
    assert rel_err(got, want) <= REL


def test_hybrid_fusion_dropout_vs_oracle(pkg, dev, world, golden):
    m = make_model(pkg, dev, world)
    fm = m.fusion_model
    pkg.MCDropoutUncertainty._set_dropout_train(fm)
    z = {"audio": torch.from_numpy(golden["z_audio"]), "video": torch.from_numpy(golden["z_video"])}
    zd = {k: v.to(dev) for k, v in z.items()}
    z64 = {k: v.double() for k, v in z.items()}
    ids = np.arange(6, dtype=np.uint64) + np.uint64(100)
    mask = _masks(world)["mixed"]
    for t in (0, 3):
        with pkg.philox_context(SEED, t, 100):
            got, info = fm(zd, torch.from_numpy(mask).to(dev), return_attention=True)
        ctx = om.DropCtx(seed=SEED, t=t, clip_ids=ids, p_fusion=0.3)
        want, winfo = om.hybrid_fusion(z64, torch.from_numpy(mask), world["sd64"], ctx=ctx, return_info=True)
        assert rel_err(got.cpu().numpy(), want.numpy()) <= REL
        assert rel_err(info["per_modality_attention"]["video"].cpu().numpy(), winfo["per_modality_attention"][1].numpy()) <= REL
        assert rel_err(info["fusion_weights"].cpu().numpy(), winfo["fusion_weights"].numpy()) <= REL


def test_cross_modal_attention_general_shapes(pkg, dev, golden):
    att = pkg.CrossModalAttention(query_dim=512, key_dim=64, hidden_dim=256, num_heads=4, dropout=0.1)
    att.load_state_dict({k[len("xattn.att."):]: torch.from_numpy(golden[k]) for k in golden.files
                         if k.startswith("xattn.att.")}, strict=True)
    att = att.to(dev).eval()
    q, kv = torch.from_numpy(golden["xattn.q"]).to(dev), torch.from_numpy(golden["xattn.kv"]).to(dev)
    inv = torch.from_numpy(golden["xattn.invalid"]).to(dev)
    out, attn = att(q, kv, kv, inv)
    assert out.shape == (4, 3, 256) and attn.shape == (4, 4, 3, 5)
    assert rel_err(out.cpu().numpy(), golden["xattn.out"]) <= REL
    assert rel_err(attn.cpu().numpy(), golden["xattn.attn"]) <= REL
    out1, attn1 = att(q[:, 0], kv, kv, (~inv).float())             # (B, D) query, numeric 1 = valid mask
    assert out1.shape == (4, 256) and rel_err(out1.cpu().numpy(), golden["xattn.out"][:, 0]) <= REL
    # dropout on the attention weights draws the spec masks
    att.attn_dropout.train()
    with pkg.philox_context(SEED, 2, 9):
        _, attn_d = att(q, kv, kv, inv)
    keep = px.dropout_keep(4 * 3 * 5, np.arange(4, dtype=np.uint64) + np.uint64(9), 2, px.SITE_ATTN, 0.1, SEED)
    want = golden["xattn.attn"].reshape(4, -1) * keep / (1 - 0.1)
    assert rel_err(attn_d.cpu().numpy().reshape(4, -1), want) <= REL


# ---------------------------------------------------------------------------------------------------------
# whole path + a8 MC-dropout
# ---------------------------------------------------------------------------------------------------------
def _features(world, dev, u8=False):
    v = torch.from_numpy(world["frames_u8"] if u8 else world["frames"]).to(dev)
    return {"audio": torch.from_numpy(world["wave"]).to(dev), "video": v}


@pytest.mark.parametrize("head", ["hybrid", "concat", "early", "late"])
def test_full_path_vs_reference_golden(pkg, dev, world, golden, head):
    m = make_model(pkg, dev, world, head)
    feats = _features(world, dev)
    for name, mask in _masks(world).items():
        if head == "concat" and name != "ones":
            continue
        got = m(feats, torch.from_numpy(mask).to(dev)).cpu().numpy()
        want = golden[f"logits_{head}_{name}"]
        assert rel_err(got, want) <= REL
        assert (got.argmax(1) == want.argmax(1)).all()
    got8 = m(_features(world, dev, u8=True)).cpu().numpy()
    assert rel_err(got8, golden[f"logits_{head}_ones"]) <= REL
    pre = m({"audio": torch.from_numpy(world["feats"]).to(dev), "video": feats["video"]}).cpu().numpy()
    assert rel_err(pre, golden[f"logits_{head}_ones"]) <= REL      # precomputed (B, 301, 104) features accepted


@pytest.mark.parametrize("head,names", [("hybrid", ("ones", "p05")), ("concat", ("ones",)),
                                        ("early", ("ones", "p05")), ("late", ("ones", "p05"))])
def test_mc_dropout_vs_reference_golden(pkg, dev, world, golden, head, names):
    """The reference's MCDropoutUncertainty ran unmodified over mask-injected reference modules (make_golden.py);
    clip ids there are NOT contiguous, so each clip runs with its own clip0."""
    m = make_model(pkg, dev, world, head)
    T = int(golden["t_mc"])
    ids = world["ids"]
    feats = _features(world, dev)
    for name in names:
        mask = torch.from_numpy(_masks(world)[name]).to(dev)
        means, uncs = [], []
        for b, g in enumerate(ids):
            mc = pkg.MCDropoutUncertainty(m, num_samples=T)
            mc.seed, mc.clip0 = SEED, int(g)
            mean, unc = mc({k: v[b:b + 1] for k, v in feats.items()}, mask[b:b + 1])
            means.append(mean)
            uncs.append(unc)
        mean = torch.cat(means).cpu().numpy()
        unc = torch.cat(uncs).cpu().numpy()
        assert rel_err(mean, golden[f"mc_mean_logits_{head}_{name}"]) <= REL
        assert rel_err(unc, golden[f"mc_uncertainty_{head}_{name}"]) <= 1e-3
        assert (mean.argmax(1) == golden[f"mc_labels_{head}_{name}"]).all()
        assert not m.training and not m.encoders["video"]._dropout.training   # modes restored


def test_mc_fused_equals_pass_by_pass(pkg, dev, world):
    """Hoisting the pass-invariant stages changes nothing: the fused T-pass path equals T separate forwards."""
    m = make_model(pkg, dev, world)
    feats = _features(world, dev)
    mask = torch.from_numpy(_masks(world)["p05"]).to(dev)
    T = 6
    pkg.MCDropoutUncertainty._set_dropout_train(m)
    mean, unc, labels, packed = m.mc_forward(feats, mask, num_samples=T, seed=SEED, clip0=500)
    stack = []
    for t in range(T):
        with pkg.philox_context(SEED, t, 500):
            stack.append(m(feats, mask))
    stack = torch.stack(stack, dim=1)
    mean2, unc2, labels2, packed2 = pkg.mc_reduce(stack)
    assert (mean == mean2).all() and (unc == unc2).all() and (labels == labels2).all() and (packed == packed2).all()
    want_mean, want_unc = om.mc_statistics(stack.transpose(0, 1).cpu().double())
    assert rel_err(mean.cpu().numpy(), want_mean.numpy()) <= 1e-6
    assert rel_err(unc.cpu().numpy(), want_unc.numpy()) <= 1e-4
    assert (labels.cpu().numpy() == want_mean.argmax(1).numpy()).all()
    assert (packed[:, :8] == mean).all() and (packed[:, 8] == unc).all()
    m.eval()
    assert float(m.mc_forward(feats, mask, num_samples=3)[1].abs().max()) < 1e-12   # eval: identical passes


def test_shard_invariance_and_determinism(pkg, dev, world):
    """Size-independent properties at a benchmark-like batch: results do not depend on how clips are batched or
    sharded (masks and inputs are keyed by global clip index), and repeat bit-exactly."""
    m = make_model(pkg, dev, world)
    pkg.MCDropoutUncertainty._set_dropout_train(m)
    B, T, clip0 = 96, 8, 10_000
    def run(lo, hi):
        n = hi - lo
        f = {"audio": pkg.synth.synth_audio(n, clip0 + lo, SEED, dev), "video": pkg.synth.synth_video(n, clip0 + lo, SEED, dev)}
        mask = pkg.synth.modality_mask(n, clip0 + lo, 0.05, SEED, dev)
        return m.mc_forward(f, mask, num_samples=T, seed=SEED, clip0=clip0 + lo)[3]
    whole = run(0, B)
    again = run(0, B)
    assert (whole == again).all()
    parts = torch.cat([run(0, 40), run(40, 41), run(41, B)])
    # the LSTM / fusion tile kernels reduce in a fixed order per clip, so batching does not change a single bit
    assert (parts == whole).all()
    assert torch.isfinite(whole).all() and (whole[:, 8] >= 0).all()


# ---------------------------------------------------------------------------------------------------------
# EarlyFusion / LateFusion modules and the calibration metrics
# ---------------------------------------------------------------------------------------------------------
def test_early_late_modules_vs_reference_golden(pkg, dev, world, golden):
    z = {"audio": torch.from_numpy(golden["z_audio"]).to(dev), "video": torch.from_numpy(golden["z_video"]).to(dev)}
    dims = {"audio": 128, "video": 128}
    early = pkg.build_fusion_model("early", dims, 8, hidden_dim=256, num_heads=4, dropout=0.3)
    late = pkg.build_fusion_model("late", dims, 8, hidden_dim=256, num_heads=4, dropout=0.3)
    sd = world["sd"]
    early.load_state_dict({k[len("early_model."):]: v for k, v in sd.items() if k.startswith("early_model.")}, strict=True)
    late.load_state_dict({k[len("late_model."):]: v for k, v in sd.items() if k.startswith("late_model.")}, strict=True)
    early, late = early.to(dev).eval(), late.to(dev).eval()
    for name, mask in _masks(world).items():
        mk = torch.from_numpy(mask).to(dev)
        assert rel_err(early(z, mk).cpu().numpy(), golden[f"logits_early_{name}"]) <= REL
        lg, per = late(z, mk)
        assert rel_err(lg.cpu().numpy(), golden[f"logits_late_{name}"]) <= REL
        if name == "mixed":
            assert rel_err(per["audio"].cpu().numpy(), golden["late_per_audio_mixed"]) <= REL
            assert rel_err(per["video"].cpu().numpy(), golden["late_per_video_mixed"]) <= REL
    # no mask; a modality that is not given at all: zeros (early, src/fusion.py:120-121) / zero logits (late, :222-225)
    sdo = world["sd"]
    zc = {k: v.cpu() for k, v in z.items()}
    want = om.early_fusion(zc, None, sdo, prefix="early_model.")
    assert rel_err(early(z).cpu().numpy(), want.numpy()) <= REL
    want = om.early_fusion({"audio": zc["audio"], "video": torch.zeros_like(zc["video"])}, None, sdo, prefix="early_model.")
    assert rel_err(early({"audio": z["audio"]}).cpu().numpy(), want.numpy()) <= REL
    lg, per = late({"video": z["video"]})
    wl, wper = om.late_fusion(zc, None, sdo, prefix="late_model.")
    assert float(per["audio"].abs().max()) == 0.0
    assert rel_err(per["video"].cpu().numpy(), wper[1].numpy()) <= REL
    w = torch.softmax(sdo["late_model.fusion_logits"], dim=0)
    assert rel_err(lg.cpu().numpy(), (w[1] * wper[1]).numpy()) <= REL
    with pytest.raises(ValueError):
        early({"imu": z["audio"]})
    with pytest.raises(ValueError):
        early(z, torch.ones(6, 3, device=dev))
    with pytest.raises(ValueError):
        pkg.build_fusion_model("mid", dims, 8)


def test_early_late_dropout_vs_oracle(pkg, dev, world, golden):
    """Dropout sites EARLY0 / EARLY1 / LATE+m draw the spec masks: compare one stochastic pass with the oracle."""
    z = {"audio": torch.from_numpy(golden["z_audio"]), "video": torch.from_numpy(golden["z_video"])}
    zd = {k: v.to(dev) for k, v in z.items()}
    dims = {"audio": 128, "video": 128}
    ids = world["ids"]
    mask = torch.from_numpy(_masks(world)["mixed"])
    for head, fn in (("early", om.early_fusion), ("late", om.late_fusion)):
        mod = pkg.build_fusion_model(head, dims, 8, hidden_dim=256, dropout=0.3)
        mod.load_state_dict({k[len(head) + 7:]: v for k, v in world["sd"].items() if k.startswith(head + "_model.")},
                            strict=True)
        mod = mod.to(dev).train()
        for b, g in enumerate(ids):
            ctx = om.DropCtx(SEED, 3, np.asarray([g], dtype=np.uint64), p_fusion=0.3)
            want = fn({k: v[b:b + 1] for k, v in z.items()}, mask[b:b + 1], world["sd"], prefix=head + "_model.", ctx=ctx)
            want = want[0] if isinstance(want, tuple) else want
            with pkg.philox_context(SEED, 3, int(g)):
                got = mod({k: v[b:b + 1] for k, v in zd.items()}, mask[b:b + 1].to(dev))
            got = got[0] if isinstance(got, tuple) else got
            assert rel_err(got.cpu().numpy(), want.numpy()) <= REL, (head, b)


def test_calibration_metrics_vs_reference_golden(pkg, dev, golden):
    logits = torch.from_numpy(golden["cal.logits"]).to(dev)
    labels = torch.from_numpy(golden["cal.labels"]).to(dev)
    probs = torch.softmax(logits, dim=1)
    confs, preds = probs.max(dim=1)
    CM = pkg.CalibrationMetrics
    assert abs(CM.expected_calibration_error(confs, preds, labels, 15) - float(golden["cal.ece"])) <= 1e-6
    assert abs(CM.maximum_calibration_error(confs, preds, labels, 15) - float(golden["cal.mce"])) <= 1e-6
    assert abs(CM.expected_calibration_error(confs, preds, labels, 10) - float(golden["cal.ece10"])) <= 1e-6
    assert abs(CM.negative_log_likelihood(logits, labels) - float(golden["cal.nll"])) <= 1e-5

    class Ident(torch.nn.Module):
        def forward(self, x):
            return x, None                                                     # (logits, aux) is unwrapped

    loader = [(logits[i:i + 1000], labels[i:i + 1000]) for i in range(0, 4000, 1000)]
    got = pkg.compute_calibration_metrics(Ident(), loader, device=dev, num_bins=15)
    for k in ("ece", "mce", "nll", "accuracy"):
        assert abs(got[k] - float(golden["cal." + k])) <= 1e-5, k
    # size-independent property: accumulating in any batching gives the same histogram
    a1 = pkg.uncertainty.calibration_accumulate(torch.zeros(48, device=dev, dtype=torch.float64), labels, 15, logits=logits)
    a2 = torch.zeros(48, device=dev, dtype=torch.float64)
    for i in range(0, 4000, 37):
        pkg.uncertainty.calibration_accumulate(a2, labels[i:i + 37], 15, logits=logits[i:i + 37])
    assert (a1[:15] == a2[:15]).all() and (a1[30:45] == a2[30:45]).all() and a1[47] == a2[47] == 4000
    assert torch.allclose(a1, a2, rtol=1e-12, atol=1e-9)


def test_gru_sequence_encoder_vs_reference_golden(pkg, dev, world, golden):
    """SequenceEncoder(encoder_type='gru') (src/encoders.py:65-77, :137-155)."""
    enc = pkg.SequenceEncoder(input_dim=104, hidden_dim=256, output_dim=128, num_layers=2, encoder_type="gru")
    enc.load_state_dict({k[len("gru_enc."):]: v for k, v in world["sd"].items() if k.startswith("gru_enc.")}, strict=True)
    enc = enc.to(dev).eval()
    feats = torch.from_numpy(world["feats"]).to(dev)
    assert rel_err(enc(feats).cpu().numpy(), golden["gru.z"]) <= REL
    lengths = torch.from_numpy(golden["gru.lengths"])
    assert rel_err(enc(feats, lengths).cpu().numpy(), golden["gru.z_lengths"]) <= REL
    # a batch that spans several recurrence tiles, ragged lengths, against the fp64 oracle
    rng = np.random.default_rng(3)
    x = torch.from_numpy(rng.normal(size=(70, 23, 104)).astype(np.float32))
    ln = torch.from_numpy(rng.integers(1, 24, size=70))
    sd64 = {k: v.double() for k, v in world["sd"].items() if k.startswith("gru_enc.")}
    want = om.sequence_encoder_gru(x.double(), sd64, lengths=ln)
    assert rel_err(enc(x.to(dev), ln).cpu().numpy(), want.numpy()) <= REL
    with pytest.raises(ValueError):
        pkg.SequenceEncoder(input_dim=104, encoder_type="conformer")


def test_cnn_sequence_encoder_vs_reference_golden(pkg, dev, world, golden):
    """SequenceEncoder(encoder_type='cnn') (src/encoders.py:79-94, :157-168): eval output, the
    stochastic passes (one nn.Dropout used twice: sites CNN0 / CNN1, clip ids of the golden file are not contiguous so
    each clip runs with its own clip0), and MCDropoutUncertainty's generic pass-by-pass path over it."""
    enc = pkg.SequenceEncoder(input_dim=104, hidden_dim=256, output_dim=128, num_layers=2, encoder_type="cnn", dropout=0.1)
    sd = {k[len("cnn_enc."):]: v for k, v in world["sd"].items() if k.startswith("cnn_enc.")}
    sd["bn1.num_batches_tracked"] = torch.tensor(0)
    sd["bn2.num_batches_tracked"] = torch.tensor(0)
    enc.load_state_dict(sd, strict=True)
    enc = enc.to(dev).eval()
    feats = torch.from_numpy(world["feats"]).to(dev)
    assert rel_err(enc(feats).cpu().numpy(), golden["cnn.z"]) <= REL
    assert rel_err(enc(feats, torch.tensor([5] * feats.shape[0])).cpu().numpy(), golden["cnn.z"]) <= REL   # lengths ignored
    pkg.MCDropoutUncertainty._set_dropout_train(enc)
    for t in range(3):
        rows = []
        for b, g in enumerate(world["ids"]):
            with pkg.philox_context(SEED, t, int(g)):
                rows.append(enc(feats[b:b + 1]))
        assert rel_err(torch.cat(rows).cpu().numpy(), golden[f"cnn.z_pass{t}"]) <= REL, t
    # several passes in one call == pass by pass; the wrapper restores modes
    many = enc._forward_cnn(feats[:1], passes=3, seed=SEED, clip0=int(world["ids"][0]), t0=0)
    assert rel_err(many[0].cpu().numpy(), np.stack([golden[f"cnn.z_pass{t}"][0] for t in range(3)])) <= REL
    enc.eval()

    class FirstEight(torch.nn.Module):               # 8 "logits": avhot_mc_reduce handles up to 32 classes
        def __init__(self, e):
            super().__init__()
            self.e = e

        def forward(self, x):
            return self.e(x)[:, :8]

    mc = pkg.MCDropoutUncertainty(FirstEight(enc).eval(), num_samples=4)
    mean, unc = mc(feats)
    assert mean.shape == (feats.shape[0], 8) and unc.shape == (feats.shape[0],) and float(unc.min()) > 0
    assert not enc.training and not enc.dropout_layer.training
    enc.train()
    with pytest.raises(ValueError):
        enc(feats)                                   # training-mode BatchNorm is not built


def test_late_fusion_with_uncertainty_vs_reference_golden(pkg, dev, world, golden):
    """build_fusion_model('uncertainty') = LateFusionWithUncertainty (src/fusion.py:589-655) and
    UncertaintyWeightedFusion (src/uncertainty.py:279-318)."""
    dims = {"audio": 128, "video": 128}
    mod = pkg.build_fusion_model("uncertainty", dims, 8, hidden_dim=256, num_heads=4, dropout=0.3)
    mod.load_state_dict({k[len("unc_model."):]: v for k, v in world["sd"].items() if k.startswith("unc_model.")}, strict=True)
    mod = mod.to(dev).eval()
    z = {"audio": torch.from_numpy(golden["z_audio"]).to(dev), "video": torch.from_numpy(golden["z_video"]).to(dev)}
    masks = _masks(world)
    for name in ("ones", "mixed"):
        lg, info = mod(z, torch.from_numpy(masks[name]).to(dev))
        assert rel_err(lg.cpu().numpy(), golden[f"unc.logits_{name}"]) <= REL
        assert rel_err(info["per_modality_logits"].cpu().numpy(), golden[f"unc.per_{name}"]) <= REL
        assert rel_err(info["fusion_weights"].cpu().numpy(), golden[f"unc.weights_{name}"]) <= REL
    lg, _ = mod({"video": z["video"]}, torch.from_numpy(masks["mixed"]).to(dev))
    assert rel_err(lg.cpu().numpy(), golden["unc.logits_video_only"]) <= REL
    # one stochastic pass: eight dropout sites, clip ids of the golden file are not contiguous
    mod.train()
    rows, wrows = [], []
    mk = torch.from_numpy(masks["mixed"]).to(dev)
    for b, g in enumerate(world["ids"]):
        with pkg.philox_context(SEED, 2, int(g)):
            lg, info = mod({k: v[b:b + 1] for k, v in z.items()}, mk[b:b + 1])
        rows.append(lg)
        wrows.append(info["fusion_weights"])
    assert rel_err(torch.cat(rows).cpu().numpy(), golden["unc.logits_drop"]) <= REL
    assert rel_err(torch.cat(wrows).cpu().numpy(), golden["unc.weights_drop"]) <= REL
    # the fuser on its own, and the hidden_dim = 0 variant (a single Linear per head) against torch arithmetic
    fuser = pkg.UncertaintyWeightedFusion(epsilon=1e-6)
    per = torch.from_numpy(golden["unc.per_mixed"]).to(dev)
    u = torch.rand(per.shape[0], 2, device=dev) + 0.1
    fused, w = fuser({"audio": per[:, 0], "video": per[:, 1]}, {"audio": u[:, 0], "video": u[:, 1]}, mk)
    inv = (1.0 / (u + 1e-6)) * mk
    want_w = inv / (inv.sum(1, keepdim=True) + 1e-6)
    assert rel_err(w.cpu().numpy(), want_w.cpu().numpy()) <= 1e-6
    assert rel_err(fused.cpu().numpy(), (want_w.unsqueeze(-1) * per).sum(1).cpu().numpy()) <= 1e-6
    lin = pkg.LateFusionWithUncertainty(dims, 8).to(dev).eval()                      # hidden_dim = 0 (the ctor default)
    lg, info = lin(z, mk)
    logits = torch.stack([z[m] @ lin.classifiers[m][1].weight.T + lin.classifiers[m][1].bias for m in dims], dim=1)
    uu = torch.stack([torch.nn.functional.softplus(z[m] @ lin.uncert_heads[m][1].weight.T + lin.uncert_heads[m][1].bias).squeeze(-1)
                      for m in dims], dim=1)
    inv = (1.0 / (uu + 1e-6)) * mk
    ww = inv / (inv.sum(1, keepdim=True) + 1e-6)
    assert rel_err(lg.cpu().numpy(), (ww.unsqueeze(-1) * logits).sum(1).detach().cpu().numpy()) <= REL
    with pytest.raises(AssertionError):
        mod(z, None)


def test_split_reader_feeds_the_path(pkg, dev, world, tmp_path):
    """the .npy split format (src/data.py:51-84) read through pinned buffers, batches in the
    reference's (features, labels, mask) form, consumed by AVHotPath and compute_calibration_metrics."""
    n = 6
    audio = world["wave"][:n].reshape(n, -1, 1)                           # (N, 48000, 1) as load_raw_audio saves it
    video = world["frames_u8"][:n]
    labels = np.arange(n) % 8
    d = {"audio": audio, "video": video, "labels": labels}
    pkg.save_splits_to_disk(d, d, d, str(tmp_path))
    m = make_model(pkg, dev, world)
    want = m({"audio": torch.from_numpy(world["wave"][:n]).to(dev), "video": torch.from_numpy(video).to(dev)})
    got, labs = [], []
    for feats, lab, mask in pkg.SplitReader(str(tmp_path), ["audio", "video"], "val", batch_size=4, device=dev):
        assert feats["video"].dtype == torch.uint8 and feats["audio"].shape[1:] == (48000, 1) and (mask == 1).all()
        got.append(m(feats, mask))
        labs.append(lab)
    assert (torch.cat(got) == want).all() and (torch.cat(labs).cpu().numpy() == labels).all()
    tr = pkg.SplitReader(str(tmp_path), ["audio", "video"], "train", batch_size=4, modality_dropout=0.5, device=dev)
    masks = torch.cat([mk for _, _, mk in tr]).cpu().numpy()
    assert (masks == px.modality_mask(np.arange(n, dtype=np.uint64), 2, 0.5, pkg._rng.DEFAULT_SEED)).all()
    cal = pkg.compute_calibration_metrics(m, pkg.SplitReader(str(tmp_path), ["audio", "video"], "test", batch_size=4, device=dev),
                                          device=dev, num_bins=15)
    ref = om.calibration_metrics(want.cpu(), torch.from_numpy(labels), 15)
    for k in ("ece", "mce", "nll", "accuracy"):
        assert abs(cal[k] - ref[k]) <= 1e-5, k


def test_lstm_fp32_cluster_resident_kernel_equals_streaming_kernel(pkg, dev, tmp_path):
    """Small fp32 batches run the recurrence with W_hh resident in the shared memory of an 8-CTA cluster (lstm_cs.cu: h
    exchanged by st.async onto the receivers' mbarriers); larger ones stream W_hh per CTA (lstm.cu).  Same arithmetic in
    the same order: bit-identical embeddings, so a clip's fp32 result does not depend on the batch size's kernel choice.
    The streaming run happens in a child process (AVHOT_LSTM_CS=0 is read once)."""
    import os
    import subprocess
    import sys
    rng = np.random.default_rng(77)
    cases = ((1, 1, 7), (2, 8, 40), (2, 100, 13), (3, 300, 9), (2, 600, 5))       # (layers, B, T): 8 / 8 / 32 / 48 clips per cluster
    got = {}
    for i, (layers, B, T) in enumerate(cases):
        torch.manual_seed(500 + i)
        enc = pkg.SequenceEncoder(104, 256, 128, layers, "lstm").to(dev).eval()
        x = rng.normal(size=(B, T, 104)).astype(np.float32)
        lengths = rng.integers(1, T + 1, size=B)
        np.save(tmp_path / f"x{i}.npy", x)
        np.save(tmp_path / f"len{i}.npy", lengths)
        torch.save(enc.state_dict(), tmp_path / f"enc{i}.pt")
        got[i] = enc(torch.from_numpy(x).to(dev), torch.from_numpy(lengths)).cpu().numpy()
    code = (
        "import sys, numpy as np, torch\n"
        "import multimodal_emotion_detection_b200 as pkg\n"
        "d = sys.argv[1]\n"
        "for i, layers in enumerate((1, 2, 2, 3, 2)):\n"
        "    enc = pkg.SequenceEncoder(104, 256, 128, layers, 'lstm')\n"
        "    enc.load_state_dict(torch.load(d + f'/enc{i}.pt'))\n"
        "    enc = enc.to('cuda:0').eval()\n"
        "    out = enc(torch.from_numpy(np.load(d + f'/x{i}.npy')).to('cuda:0'), torch.from_numpy(np.load(d + f'/len{i}.npy')))\n"
        "    np.save(d + f'/stream{i}.npy', out.cpu().numpy())\n")
    env = dict(os.environ, AVHOT_LSTM_CS="0", PYTHONPATH=os.pathsep.join(sys.path))
    subprocess.run([sys.executable, "-c", code, str(tmp_path)], check=True, env=env, timeout=300)
    for i in range(len(cases)):
        assert (got[i] == np.load(tmp_path / f"stream{i}.npy")).all(), cases[i]
