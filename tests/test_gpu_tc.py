"""AVHOT_PREC_TC: the tcgen05 tensor-core path (TF32 time-batched GEMMs, fp16 recurrent GEMM, fp32 accumulate and
fp32 everything else) against the fp64 oracle at its STATED tolerance, and against the fp32 CUDA path for labels."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import frontend as ofe, model as om, philox as px, synth as osynth
from test_gpu_parity import SEED, make_model, rel_err, world, pkg, dev  # noqa: F401  (fixtures)

TC_REL_HID = 2e-3      # video Linear + ReLU output, TF32 operands over K = 12544
TC_REL_LSTM = 2e-3     # audio embedding after 2 x 301 recurrent steps with fp16 state / weights
TC_REL_LOGITS = 2e-3   # logits after LayerNorms and the fusion head


def test_frame_hidden_tc(pkg, dev, world):
    m = make_model(pkg, dev, world).set_precision("tc")
    frames = torch.from_numpy(world["frames"])
    hid = m.encoders["video"].hidden(frames.to(dev)).cpu().numpy()
    want = om.frame_hidden(frames.double(), world["sd64"]).numpy()
    err = rel_err(hid, want)
    print("frame hidden tc rel err", err)
    assert err <= TC_REL_HID


def test_lstm_tc_vs_oracle(pkg, dev, world, golden):
    m = make_model(pkg, dev, world).set_precision("tc")
    enc = m.encoders["audio"]
    feats = torch.from_numpy(world["feats"])
    got = enc(feats.to(dev)).cpu().numpy()
    want = om.sequence_encoder_lstm(feats.double(), world["sd64"]).numpy()
    err = rel_err(got, want)
    print("lstm tc rel err", err)
    assert err <= TC_REL_LSTM
    # several CTAs with a partial last tile, ragged lengths, and agreement with the fp32 CUDA path
    rng = np.random.default_rng(5)
    x = torch.from_numpy(rng.normal(size=(300, 20, 104)).astype(np.float32)).to(dev)
    lengths = torch.from_numpy(rng.integers(1, 21, size=300))
    tc_out = enc(x, lengths)
    enc.precision = pkg._lib.PREC_FP32
    ref_out = enc(x, lengths)
    err2 = rel_err(tc_out.cpu().numpy(), ref_out.cpu().numpy())
    print("lstm tc vs fp32 cuda (B=300, ragged) rel err", err2)
    assert err2 <= TC_REL_LSTM
    again = enc(x, lengths)
    assert (again == ref_out).all()


def test_lstm_tc_pair_kernel_equals_single_cta_kernel(pkg, dev, world, tmp_path):
    """The CTA-pair recurrence (cta_group::2, M = 256) and the single-CTA kernel (AVHOT_REC_PAIR=0, also what runs when
    the pairs of a grid cannot all be co-resident) compute the same rows: bit-identical embeddings.  The switch is read
    once per process, so the single-CTA run happens in a child process."""
    import os
    import subprocess
    import sys
    rng = np.random.default_rng(11)
    x = rng.normal(size=(300, 24, 104)).astype(np.float32)
    lengths = rng.integers(1, 25, size=300)
    np.save(tmp_path / "x.npy", x)
    np.save(tmp_path / "len.npy", lengths)
    enc = make_model(pkg, dev, world).set_precision("tc").encoders["audio"]
    torch.save(enc.state_dict(), tmp_path / "enc.pt")
    code = (
        "import sys, numpy as np, torch\n"
        "import multimodal_emotion_detection_b200 as pkg\n"
        "d = sys.argv[1]\n"
        "enc = pkg.SequenceEncoder(104, 256, 128, 2, 'lstm')\n"
        "enc.load_state_dict(torch.load(d + '/enc.pt'))\n"
        "enc = enc.to('cuda:0').eval(); enc.precision = pkg._lib.PREC_TC\n"
        "out = enc(torch.from_numpy(np.load(d + '/x.npy')).to('cuda:0'), torch.from_numpy(np.load(d + '/len.npy')))\n"
        "np.save(d + '/' + sys.argv[2] + '.npy', out.cpu().numpy())\n")
    # both runs with the fused layer kernel off (it would take this batch; lstm_fused.cu has its own test below)
    for name, pair_env in (("pair", "1"), ("single", "0")):
        env = dict(os.environ, AVHOT_LSTM_FUSED="0", AVHOT_REC_PAIR=pair_env, PYTHONPATH=os.pathsep.join(sys.path))
        subprocess.run([sys.executable, "-c", code, str(tmp_path), name], check=True, env=env, timeout=300)
    pair, single = np.load(tmp_path / "pair.npy"), np.load(tmp_path / "single.npy")
    assert (pair == single).all()


def test_lstm_fused_layer_kernel(pkg, dev, world, tmp_path):
    """lstm_fused.cu (input projection + recurrence of a layer in one kernel, h in tensor memory, c in registers) is the
    tensor-core LSTM at every batch size.  Its roles - layer 0 on the split operand with the hidden sequence stored, a
    middle layer in place, the top layer, a single layer - against the fp32 CUDA kernels at the stated tolerance:
    ragged lengths, a single tile (the pair's second CTA is a duplicate that writes nothing), odd tile counts; and
    against the two-kernel path (AVHOT_LSTM_FUSED=0, child process: the switch is read once), which rounds xp to fp16
    and so is not bit-identical but sits inside the same tolerance.  Batching invariance: rows of a sub-batch are
    bit-identical to the same rows of the whole batch."""
    import os
    import subprocess
    import sys
    rng = np.random.default_rng(23)
    cases = ((1, 257, 9), (2, 300, 33), (3, 700, 17), (2, 50, 21))
    want, fused = {}, {}
    for i, (layers, B, T) in enumerate(cases):
        torch.manual_seed(100 + layers)
        enc = pkg.SequenceEncoder(104, 256, 128, layers, "lstm").to(dev).eval()
        x = rng.normal(size=(B, T, 104)).astype(np.float32) * 20.0      # dB-sized features: the split operand matters
        lengths = rng.integers(1, T + 1, size=B)
        np.save(tmp_path / f"x{i}.npy", x)
        np.save(tmp_path / f"len{i}.npy", lengths)
        torch.save(enc.state_dict(), tmp_path / f"enc{i}.pt")
        xd, ld = torch.from_numpy(x).to(dev), torch.from_numpy(lengths)
        enc.precision = pkg._lib.PREC_TC
        fused[i] = enc(xd, ld).cpu().numpy()
        lo, hi = B // 3, B // 3 + 40
        part = enc(xd[lo:hi], ld[lo:hi]).cpu().numpy()
        assert (part == fused[i][lo:hi]).all(), "a clip's embedding depends on the batch it was computed in"
        enc.precision = pkg._lib.PREC_FP32
        want[i] = enc(xd, ld).cpu().numpy()
    code = (
        "import sys, numpy as np, torch\n"
        "import multimodal_emotion_detection_b200 as pkg\n"
        "d = sys.argv[1]\n"
        "for i, layers in enumerate((1, 2, 3, 2)):\n"
        "    enc = pkg.SequenceEncoder(104, 256, 128, layers, 'lstm')\n"
        "    enc.load_state_dict(torch.load(d + f'/enc{i}.pt'))\n"
        "    enc = enc.to('cuda:0').eval(); enc.precision = pkg._lib.PREC_TC\n"
        "    out = enc(torch.from_numpy(np.load(d + f'/x{i}.npy')).to('cuda:0'), torch.from_numpy(np.load(d + f'/len{i}.npy')))\n"
        "    np.save(d + f'/two{i}.npy', out.cpu().numpy())\n")
    env = dict(os.environ, AVHOT_LSTM_FUSED="0", PYTHONPATH=os.pathsep.join(sys.path))
    subprocess.run([sys.executable, "-c", code, str(tmp_path)], check=True, env=env, timeout=300)
    for i, (layers, B, T) in enumerate(cases):
        two = np.load(tmp_path / f"two{i}.npy")
        e_f, e_u, e_fu = rel_err(fused[i], want[i]), rel_err(two, want[i]), rel_err(fused[i], two)
        print(f"{layers} layer(s), B={B}, T={T}: fused {e_f:.3g}, two-kernel path {e_u:.3g}, fused vs two-kernel {e_fu:.3g}")
        assert e_f <= TC_REL_LSTM and e_u <= TC_REL_LSTM
        assert e_fu > 0, "the fused kernel did not run (identical to the two-kernel result)"


@pytest.mark.parametrize("B,T,layers", [(1, 1, 2), (1, 5, 1), (128, 2, 2), (129, 3, 2), (257, 1, 3), (5, 301, 2)])
def test_lstm_fused_edge_shapes(pkg, dev, B, T, layers):
    """The fused layer kernel at the edges: one clip, one step (no recurrent half at all), exactly one tile, one tile plus
    one clip, full-length sequences; lengths None / all 1 / ragged - against the fp32 CUDA kernels."""
    torch.manual_seed(B * 31 + T)
    enc = pkg.SequenceEncoder(104, 256, 128, layers, "lstm").to(dev).eval()
    rng = np.random.default_rng(B + 100 * T)
    x = torch.from_numpy(rng.normal(size=(B, T, 104)).astype(np.float32) * 10.0).to(dev)
    for lengths in (None, torch.ones(B, dtype=torch.int64), torch.from_numpy(rng.integers(1, T + 1, size=B))):
        enc.precision = pkg._lib.PREC_TC
        got = enc(x, lengths).cpu().numpy()
        enc.precision = pkg._lib.PREC_FP32
        ref = enc(x, lengths).cpu().numpy()
        assert np.isfinite(got).all()
        assert rel_err(got, ref) <= TC_REL_LSTM, (B, T, layers, rel_err(got, ref))


def test_lstm_dct_fold_on_frontend_operand(pkg, dev, world):
    """forward_from_waveform hands the LSTM the front-end's split operand, whose 40 MFCC columns are the DCT of its 64
    log-mel columns; layer 0 then multiplies the log-mel k-blocks alone by W[:, :64] + W[:, 64:] D
    (avhot_lstm_forward_frontend, K = 192 instead of 320).  Same embedding as the unfolded tensor-core path on the fp32
    features and as the fp32 kernels, at the stated tolerance; small and large batches (fold + fused layer kernel)."""
    torch.manual_seed(3)
    enc = pkg.SequenceEncoder(104, 256, 128, 2, "lstm").to(dev).eval()
    for B in (6, 300, 3000):
        wave = pkg.synth.synth_audio(B, 1000, SEED, dev, n_samples=16000)
        enc.precision = pkg._lib.PREC_TC
        folded = enc.forward_from_waveform(wave).cpu().numpy()
        feats = pkg.logmel_mfcc(wave)
        unfolded = enc(feats).cpu().numpy()
        enc.precision = pkg._lib.PREC_FP32
        ref = enc(feats).cpu().numpy()
        e1, e2 = rel_err(folded, ref), rel_err(unfolded, ref)
        print(f"B={B}: folded {e1:.3g}, unfolded tc {e2:.3g} vs fp32 kernels")
        assert e1 <= TC_REL_LSTM and e2 <= TC_REL_LSTM


def test_full_path_tc(pkg, dev, world, golden):
    m = make_model(pkg, dev, world).set_precision("tc")
    feats = {"audio": torch.from_numpy(world["wave"]).to(dev), "video": torch.from_numpy(world["frames"]).to(dev)}
    got = m(feats).cpu().numpy()
    want = golden["logits_hybrid_ones"]
    err = rel_err(got, want)
    print("full path tc logits rel err", err, "labels", got.argmax(1), want.argmax(1))
    assert err <= TC_REL_LOGITS
    # MC statistics: same masks (bit-exact Philox), tensor-core contractions
    pkg.MCDropoutUncertainty._set_dropout_train(m)
    mask = torch.from_numpy(px.modality_mask(np.arange(6, dtype=np.uint64) + np.uint64(50), 2, 0.05, SEED)).to(dev)
    mean_tc, unc_tc, lab_tc, _ = m.mc_forward(feats, mask, num_samples=8, seed=SEED, clip0=50)
    m.set_precision("fp32")
    mean, unc, lab, _ = m.mc_forward(feats, mask, num_samples=8, seed=SEED, clip0=50)
    assert rel_err(mean_tc.cpu().numpy(), mean.cpu().numpy()) <= TC_REL_LOGITS
    assert rel_err(unc_tc.cpu().numpy(), unc.cpu().numpy()) <= 5e-2


def test_hybrid_fusion_tc(pkg, dev, world, golden):
    """HybridFusion alone on the tensor-core path (fp16 GEMM operands) vs the reference golden logits and vs the
    fp32 kernels with dropout on, several 128-row tiles, partial last tile, every mask pattern."""
    m = make_model(pkg, dev, world).set_precision("tc")
    fm = m.fusion_model
    feats = {"audio": torch.from_numpy(golden["z_audio"]).to(dev), "video": torch.from_numpy(golden["z_video"]).to(dev)}
    masks = {"ones": np.ones((6, 2), np.float32),
             "mixed": np.array([[1, 1], [1, 0], [0, 1], [0, 0], [1, 1], [0, 1]], np.float32)}
    for name, mask in masks.items():
        got = fm(feats, torch.from_numpy(mask).to(dev)).cpu().numpy()
        err = rel_err(got, golden[f"logits_hybrid_{name}"])
        print("fusion tc", name, "rel err", err)
        assert err <= TC_REL_LOGITS
    # rows = clips x passes, with dropout: 300 clips x 7 passes = 2100 rows = 16.4 tiles
    rng = np.random.default_rng(11)
    B, T = 300, 7
    z_a = torch.from_numpy(rng.normal(size=(B, 128)).astype(np.float32) * 0.3).to(dev)
    z_v = torch.from_numpy(rng.normal(size=(B * T, 128)).astype(np.float32) * 0.3).to(dev)
    mask = torch.from_numpy((rng.random(size=(B, 2)) > 0.2).astype(np.float32)).to(dev)
    pkg.MCDropoutUncertainty._set_dropout_train(fm)
    p = fm._drop_p()
    assert p == 0.3
    tc_logits = fm.run([z_a, z_v], [1, T], mask, B, T, p, SEED, 1234, 5)[0]
    fm.precision = pkg._lib.PREC_FP32
    ref_logits = fm.run([z_a, z_v], [1, T], mask, B, T, p, SEED, 1234, 5)[0]
    err = rel_err(tc_logits.cpu().numpy(), ref_logits.cpu().numpy())
    print("fusion tc vs fp32 kernels (dropout, 2100 rows) rel err", err)
    assert err <= TC_REL_LOGITS
    assert torch.isfinite(tc_logits).all()


def test_fused_frontend_lstm_operand(pkg, dev, world):
    """avhot_frontend_ex writes the LSTM's split GEMM operand directly, for a batch that needs clip padding.  Layer 0 then
    reads only its log-mel k-blocks against weights with the MFCC's DCT folded in (avhot_lstm_forward_frontend), so the
    result equals running the front-end and the tensor-core LSTM separately up to the fp32 rounding of the MFCC - far
    inside the tensor-core tolerance, not bit for bit."""
    m = make_model(pkg, dev, world).set_precision("tc")
    enc = m.encoders["audio"]
    wave = pkg.synth.synth_audio(37, 900, SEED, dev)
    fused = enc.forward_from_waveform(wave)
    separate = enc(pkg.logmel_mfcc(wave))
    err = rel_err(fused.cpu().numpy(), separate.cpu().numpy())
    print("front-end operand (DCT folded) vs features path:", err)
    assert err <= 5e-4
    again = enc.forward_from_waveform(wave)
    assert (again == fused).all()


def test_u8_ingest_tc(pkg, dev, world):
    """uint8 pixels on the tensor-core path go straight onto tcgen05 (gemm_u8.cu: pixels exact in fp16, weights W s / 255
    in fp16, one pass): against the fp64 oracle at the tensor-core tolerance, and at least as close to it as the TF32
    GEMM that takes the fp32 frames; several tiles with a partial last one; deterministic."""
    m = make_model(pkg, dev, world).set_precision("tc")
    enc = m.encoders["video"]
    want = om.frame_hidden(torch.from_numpy(world["frames"]).double(), world["sd64"]).numpy()
    h32 = enc.hidden(torch.from_numpy(world["frames"]).to(dev)).cpu().numpy()
    h8 = enc.hidden(torch.from_numpy(world["frames_u8"]).to(dev)).cpu().numpy()
    e8, e32 = rel_err(h8, want), rel_err(h32, want)
    print("frame hidden tc: u8 fp16 GEMM rel err", e8, " fp32 frames (converted to fp16 in the kernel)", e32)
    assert e8 <= TC_REL_HID and e32 <= TC_REL_HID and e8 <= 1.5 * e32
    B = 21                                                   # 336 rows: two full 128-row tiles and a partial one
    frames = pkg.synth.synth_video(B, 5, SEED, dev, dtype=torch.uint8)
    got = enc.hidden(frames)
    enc.precision = pkg._lib.PREC_FP32
    ref = enc.hidden(frames)
    enc.precision = pkg._lib.PREC_TC
    assert got.shape == (B, 16, 256) and rel_err(got.cpu().numpy(), ref.cpu().numpy()) <= TC_REL_HID
    assert torch.equal(enc.hidden(frames), got)
    assert float(got.min()) >= 0.0                           # ReLU


def test_u8_pair_gemm_equals_single_cta_kernel(pkg, dev, world, tmp_path):
    """uint8 frames: the CTA-pair converting GEMM (gemm_cvt.cu, M = 256, weights shared by the pair) and its single-CTA
    predecessor (gemm_u8.cu, AVHOT_FRAME_U8_PAIR=0, child process: the switch is read once) run the same fp16 MMA chain
    per row: bit-identical hidden activations, odd tile counts and partial tiles included."""
    import os
    import subprocess
    import sys
    m = make_model(pkg, dev, world).set_precision("tc")
    enc = m.encoders["video"]
    torch.save(enc.state_dict(), tmp_path / "venc.pt")
    outs = {}
    for B in (1, 21, 300):
        frames = pkg.synth.synth_video(B, 3, SEED, dev, dtype=torch.uint8)
        outs[B] = enc.hidden(frames).cpu().numpy()
    code = (
        "import sys, numpy as np, torch\n"
        "import multimodal_emotion_detection_b200 as pkg\n"
        "d = sys.argv[1]\n"
        "enc = pkg.FrameEncoder(12544, 256, 128, 'attention', 0.1)\n"
        "enc.load_state_dict(torch.load(d + '/venc.pt'))\n"
        "enc = enc.to('cuda:0').eval(); enc.precision = pkg._lib.PREC_TC\n"
        "for B in (1, 21, 300):\n"
        "    frames = pkg.synth.synth_video(B, 3, 42, 'cuda:0', dtype=torch.uint8)\n"
        "    np.save(d + f'/single{B}.npy', enc.hidden(frames).cpu().numpy())\n")
    env = dict(os.environ, AVHOT_FRAME_U8_PAIR="0", PYTHONPATH=os.pathsep.join(sys.path))
    subprocess.run([sys.executable, "-c", code, str(tmp_path)], check=True, env=env, timeout=300)
    for B, got in outs.items():
        assert (got == np.load(tmp_path / f"single{B}.npy")).all(), B


def test_fp32_frames_cvt_pair_gemm(pkg, dev, world):
    """fp32 frames on the tensor-core path: gemm_cvt.cu converts them to fp16 inside a CTA-pair GEMM (M = 256 per pair,
    fp16 weights shared).  Odd numbers of 128-row tiles (the pair's second CTA has no rows), partial tiles, more tiles
    than pairs on the GPU, against the fp32 kernels at the stated tolerance; rows do not depend on the batch they are
    computed in; deterministic."""
    m = make_model(pkg, dev, world).set_precision("tc")
    enc = m.encoders["video"]
    for B in (1, 21, 40, 1250):                               # 16 / 336 / 640 / 20000 rows
        frames = pkg.synth.synth_video(B, 11, SEED, dev)
        enc.precision = pkg._lib.PREC_TC
        got = enc.hidden(frames)
        assert torch.equal(enc.hidden(frames), got)
        lo = B // 2
        assert torch.equal(enc.hidden(frames[lo:lo + 3]), got[lo:lo + 3])
        enc.precision = pkg._lib.PREC_FP32
        ref = enc.hidden(frames)
        err = rel_err(got.cpu().numpy(), ref.cpu().numpy())
        print(f"B={B}: fp16-converted pair GEMM vs fp32 kernels {err:.3g}")
        assert err <= TC_REL_HID and float(got.min()) >= 0.0
    enc.precision = pkg._lib.PREC_TC


@pytest.mark.parametrize("B", [1, 129])
def test_tc_edge_batches(pkg, dev, world, B):
    """One clip, and 128 + 1 clips (a second 128-clip tile with a single live row): tensor-core path == fp32 path
    within tolerance for the whole MC pipeline, labels equal up to near-ties."""
    m = make_model(pkg, dev, world)
    pkg.MCDropoutUncertainty._set_dropout_train(m)
    feats = {"audio": pkg.synth.synth_audio(B, 77, SEED, dev), "video": pkg.synth.synth_video(B, 77, SEED, dev)}
    mask = pkg.synth.modality_mask(B, 77, 0.05, SEED, dev)
    ref = m.mc_forward(feats, mask, num_samples=4, seed=SEED, clip0=77)
    m.set_precision("tc")
    got = m.mc_forward(feats, mask, num_samples=4, seed=SEED, clip0=77)
    assert rel_err(got[0].cpu().numpy(), ref[0].cpu().numpy()) <= TC_REL_LOGITS
    assert rel_err(got[1].cpu().numpy(), ref[1].cpu().numpy()) <= 5e-2
    assert torch.isfinite(got[3]).all()
    assert (got[2] != ref[2]).sum().item() <= max(1, B // 100)


def test_folded_video_projection(pkg, dev, world):
    """The tensor-core hybrid path folds FrameEncoder.projection[1] into HybridFusion.proj_layers.video (two linear maps
    back to back) and hands the fusion kernel the LayerNorm output instead of z_v: same masks, same result up to the
    rounding of the folded fp16 weight; the unfolded pipeline stays selectable."""
    m = make_model(pkg, dev, world).set_precision("tc")
    B, T, clip0 = 300, 6, 4242
    feats = {"audio": pkg.synth.synth_audio(B, clip0, SEED, dev), "video": pkg.synth.synth_video(B, clip0, SEED, dev)}
    mask = pkg.synth.modality_mask(B, clip0, 0.05, SEED, dev)
    pkg.MCDropoutUncertainty._set_dropout_train(m)
    assert m.fold_video_projection
    folded = m.mc_forward(feats, mask, num_samples=T, seed=SEED, clip0=clip0)
    m.fold_video_projection = False
    plain = m.mc_forward(feats, mask, num_samples=T, seed=SEED, clip0=clip0)
    m.set_precision("fp32")
    exact = m.mc_forward(feats, mask, num_samples=T, seed=SEED, clip0=clip0)
    e_fold, e_plain = rel_err(folded[0].cpu().numpy(), exact[0].cpu().numpy()), rel_err(plain[0].cpu().numpy(), exact[0].cpu().numpy())
    print("folded vs fp32", e_fold, "unfolded vs fp32", e_plain, "folded vs unfolded",
          rel_err(folded[0].cpu().numpy(), plain[0].cpu().numpy()))
    assert e_fold <= TC_REL_LOGITS and e_plain <= TC_REL_LOGITS
    assert rel_err(folded[1].cpu().numpy(), exact[1].cpu().numpy()) <= 5e-2
    # the LayerNorm rows themselves (avhot_frame_tail_norm) against the oracle, one stochastic pass
    venc = m.encoders["video"]
    hid = venc.hidden(feats["video"][:4])
    y = venc.tail(hid, None, 2, 0.1, SEED, clip0, 0, project=False)
    z = venc.tail(hid, None, 2, 0.1, SEED, clip0, 0, project=True)
    lin = venc.projection[1]
    assert y.shape == (4, 2, 256)
    want = torch.nn.functional.linear(y.double(), lin.weight.double(), lin.bias.double())
    assert rel_err(z.cpu().numpy(), want.detach().cpu().numpy()) <= 1e-5
    # a changed projection weight rebuilds the folded handle
    m.set_precision("tc")
    m.fold_video_projection = True
    with torch.no_grad():
        lin.weight.mul_(1.5)
    again = m.mc_forward(feats, mask, num_samples=T, seed=SEED, clip0=clip0)
    assert rel_err(again[0].cpu().numpy(), folded[0].cpu().numpy()) > 1e-3


def test_fusion_tc_column_halves_equal_whole_rows(pkg, dev, world, tmp_path):
    """The eight-epilogue-warp fusion kernel (two warps share a row and split its columns) against the four-warp form
    (thread = whole row, AVHOT_FUSION_HALVES=1, read once per process -> child process): same masks, same arithmetic
    up to the order in which the parts of a LayerNorm / classifier sum are added (which can flip the last bit of an
    fp16 operand of the next GEMM: 5e-4 relative on one element)."""
    import os
    import subprocess
    import sys
    fm = make_model(pkg, dev, world).set_precision("tc").fusion_model
    torch.save(fm.state_dict(), tmp_path / "fm.pt")
    rng = np.random.default_rng(21)
    B, T = 37, 9                                           # 333 rows: two full tiles and a partial one
    za = rng.normal(size=(B, 128)).astype(np.float32)
    zv = rng.normal(size=(B * T, 128)).astype(np.float32)
    mask = (rng.uniform(size=(B, 2)) > 0.2).astype(np.float32)
    np.save(tmp_path / "za.npy", za); np.save(tmp_path / "zv.npy", zv); np.save(tmp_path / "mask.npy", mask)
    code = (
        "import sys, numpy as np, torch\n"
        "import multimodal_emotion_detection_b200 as pkg\n"
        "d, out = sys.argv[1], sys.argv[2]\n"
        "fm = pkg.build_fusion_model('hybrid', {'audio': 128, 'video': 128}, 8, hidden_dim=256, num_heads=4, dropout=0.3)\n"
        "fm.load_state_dict(torch.load(d + '/fm.pt'))\n"
        "fm = fm.to('cuda:0').eval(); fm.precision = pkg._lib.PREC_TC\n"
        "L = lambda n: torch.from_numpy(np.load(d + '/' + n + '.npy')).to('cuda:0')\n"
        "lg = fm.run([L('za'), L('zv')], [1, 9], L('mask'), 37, 9, 0.3, 42, 777, 0)[0]\n"
        "np.save(d + '/' + out, lg.cpu().numpy())\n")
    for halves, out in (("1", "rows.npy"), ("2", "halves.npy"), ("4", "quarters.npy")):
        env = dict(os.environ, AVHOT_FUSION_HALVES=halves, PYTHONPATH=os.pathsep.join(sys.path))
        subprocess.run([sys.executable, "-c", code, str(tmp_path), out], check=True, env=env, timeout=300)
    rows = np.load(tmp_path / "rows.npy")
    assert rows.shape == (B, T, 8) and np.isfinite(rows).all()
    for name in ("halves.npy", "quarters.npy"):
        err = rel_err(np.load(tmp_path / name), rows)
        print("fusion_tc", name, "vs rows", err)
        assert err <= 3e-4
