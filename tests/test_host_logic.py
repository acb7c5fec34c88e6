"""Host-side mirror of the reference interface: signatures, state-dict layout, error
behaviour, and the no-CPU-fallback rule.  No GPU needed."""
import inspect

import numpy as np
import pytest
import torch

import multimodal_emotion_detection_b200 as pkg
from oracle import synth


def test_signatures_match_reference():
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(pkg.SequenceEncoder.__init__)[1:] == ["input_dim", "hidden_dim", "output_dim", "num_layers",
                                                     "encoder_type", "dropout"]          # src/encoders.py:23-31
    assert sig(pkg.SequenceEncoder.forward)[1:] == ["sequence", "lengths"]
    assert sig(pkg.FrameEncoder.__init__)[1:] == ["frame_dim", "hidden_dim", "output_dim", "temporal_pooling",
                                                  "dropout"]                              # :203-210
    assert sig(pkg.FrameEncoder.forward)[1:] == ["frames", "mask"]
    assert sig(pkg.build_encoder) == ["modality", "input_dim", "output_dim", "encoder_config"]
    assert sig(pkg.CrossModalAttention.__init__)[1:] == ["query_dim", "key_dim", "hidden_dim", "num_heads",
                                                         "dropout"]                       # src/attention.py:31-32
    assert sig(pkg.CrossModalAttention.forward)[1:] == ["query", "key", "value", "mask"]
    assert sig(pkg.HybridFusion.__init__)[1:] == ["modality_dims", "hidden_dim", "num_classes", "num_heads",
                                                  "dropout"]                              # src/fusion.py:409-416
    assert sig(pkg.HybridFusion.forward)[1:] == ["modality_features", "modality_mask", "return_attention"]
    assert sig(pkg.build_fusion_model)[:3] == ["fusion_type", "modality_dims", "num_classes"]
    assert sig(pkg.MCDropoutUncertainty.__init__)[1:] == ["model", "num_samples"]         # src/uncertainty.py:25
    assert inspect.signature(pkg.HybridFusion.__init__).parameters["num_classes"].default == 11
    assert inspect.signature(pkg.MCDropoutUncertainty.__init__).parameters["num_samples"].default == 10
    assert sig(pkg.load_raw_audio) == ["wav_path", "sr", "max_duration"]                  # src/dataprocessing.py:184-188
    # SURVEY.md 8(f) row 4
    assert sig(pkg.TemporalAttention.__init__)[1:] == ["feature_dim", "hidden_dim", "num_heads", "dropout"]   # src/attention.py:150-156
    assert sig(pkg.TemporalAttention.forward)[1:] == ["sequence", "mask"]                  # :181-185
    assert sig(pkg.TemporalAttention.pool_sequence)[1:] == ["sequence", "attention_weights"]   # :245-249
    assert sig(pkg.PairwiseModalityAttention.__init__)[1:] == ["modality_dims", "hidden_dim", "num_heads", "dropout"]   # :427-433
    assert sig(pkg.PairwiseModalityAttention.forward)[1:] == ["modality_features", "modality_mask"]           # :470-474
    assert inspect.signature(pkg.TemporalAttention.__init__).parameters["hidden_dim"].default == 256


def test_sequence_attention_modules_parameter_layout():
    """Parameter names / shapes of the 8(f) modules as the reference's classes create them (counts from the reference:
    SequenceEncoder(transformer, 104 -> 256 -> 128, 2 layers) 2 687 872 incl. the 4096-row position table;
    TemporalAttention(104, 256) 146 432; PairwiseModalityAttention({a: 128, v: 128}, 256) 397 824)."""
    enc = pkg.SequenceEncoder(input_dim=104, hidden_dim=256, output_dim=128, num_layers=2, encoder_type="transformer")
    sd = enc.state_dict()
    assert sd["pos_embedding.weight"].shape == (4096, 256) and sd["input_proj.weight"].shape == (256, 104)
    assert sd["transformer.layers.1.self_attn.in_proj_weight"].shape == (768, 256)
    assert sd["transformer.layers.0.linear1.weight"].shape == (1024, 256) and sd["projection.weight"].shape == (128, 256)
    assert sum(p.numel() for p in enc.parameters()) == 2_687_872
    ta = pkg.TemporalAttention(104, 256)
    assert sorted(ta.state_dict()) == sorted(f"{n}.{w}" for n in ("q_proj", "k_proj", "v_proj", "out_proj") for w in ("weight", "bias"))
    assert sum(p.numel() for p in ta.parameters()) == 3 * (104 * 256 + 256) + 256 * 256 + 256 == 146_432
    pm = pkg.PairwiseModalityAttention({"audio": 128, "video": 128}, hidden_dim=256)
    assert set(k.split(".")[0] for k in pm.state_dict()) == {"cross_attn", "self_proj", "out_ln"}
    per_dir = 3 * 2 * 128 + 3 * (128 * 256 + 256) + 256 * 256 + 256          # three input LayerNorms, q / k / v, out_proj
    assert sum(p.numel() for p in pm.parameters()) == 2 * per_dir + 2 * (128 * 256 + 256) + 2 * 256 == 397_824
    with pytest.raises(ValueError):
        pkg.PairwiseModalityAttention({str(i): 8 for i in range(5)})


def test_state_dict_layout_appendix_a():
    want = synth.model_shapes()
    for head in ("hybrid", "concat", "early", "late"):
        sd = pkg.AVHotPath(head).state_dict()
        exp = synth.state_dict_for(head, {k: v[0] for k, v in want.items()})
        assert set(sd) == set(exp), head
        for k, shape in exp.items():
            assert tuple(sd[k].shape) == tuple(shape), k
    m = pkg.AVHotPath("hybrid")
    w = {k: torch.from_numpy(v) for k, v in synth.state_dict_for("hybrid", synth.synth_weights(42)).items()}
    m.load_state_dict(w, strict=True)
    assert sum(p.numel() for p in m.encoders["audio"].parameters()) == 929920
    assert sum(p.numel() for p in m.encoders["video"].parameters()) == 3245185
    assert sum(p.numel() for p in m.fusion_model.parameters()) == 631561


def test_error_behaviour_like_reference():
    with pytest.raises(ValueError):
        pkg.SequenceEncoder(8, encoder_type="bogus")                                      # src/encoders.py:115
    with pytest.raises(ValueError):
        pkg.FrameEncoder(8, temporal_pooling="bogus")                                     # :242
    with pytest.raises(ValueError):
        pkg.build_encoder("x", 8, 8, {"type": "bogus"})                                   # :636
    with pytest.raises(ValueError):
        pkg.build_fusion_model("bogus", {"a": 8}, 3)                                      # src/fusion.py:696
    with pytest.raises(AssertionError):
        pkg.CrossModalAttention(8, 8, 10, num_heads=4)                                    # src/attention.py:34
    with pytest.raises(AssertionError):
        pkg.MCDropoutUncertainty(torch.nn.Identity(), num_samples=0)                      # src/uncertainty.py:31
    att = pkg.CrossModalAttention(8, 8, 8, num_heads=2)
    with pytest.raises(ValueError):
        att._normalize_mask(torch.zeros(2, 2, 2), 2, 2)                                   # src/attention.py:76
    assert att._normalize_mask(torch.tensor([1.0, 0.0]), 2, 3).tolist() == [[False] * 3, [True] * 3]
    assert isinstance(pkg.build_encoder("video", 12544, 128, {"hidden_dim": 256}), pkg.FrameEncoder)
    assert isinstance(pkg.build_encoder("audio", 104, 128, {"hidden_dim": 256}), pkg.SequenceEncoder)


def test_no_cpu_fallback():
    m = pkg.AVHotPath("hybrid")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m({"audio": torch.zeros(1, 48000), "video": torch.zeros(1, 16, 12544)})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.logmel_mfcc(torch.zeros(1, 48000))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.HybridFusion({"audio": 128, "video": 128}, num_classes=8)({"audio": torch.zeros(1, 128),
                                                                       "video": torch.zeros(1, 128)})


def test_mc_wrapper_mode_handling():
    m = pkg.AVHotPath("hybrid").train()
    mc = pkg.MCDropoutUncertainty(m, num_samples=2)
    with pytest.raises(RuntimeError):
        mc({"audio": torch.zeros(1, 48000), "video": torch.zeros(1, 16, 12544)})
    assert m.training                                     # restored (src/uncertainty.py:67-71)
    m.eval()
    pkg.MCDropoutUncertainty._set_dropout_train(m)
    on = [n for n, x in m.named_modules() if isinstance(x, torch.nn.Dropout) and x.training]
    assert sorted(on) == sorted(["encoders.video.frame_mlp.2", "encoders.video._dropout",
                                 "fusion_model.attn_layers.audio.attn_dropout",
                                 "fusion_model.attn_layers.video.attn_dropout", "fusion_model.fusion_gate.2"])
    assert not m.encoders["audio"].rnn.training           # LSTM dropout stays off


def test_shard_ranges():
    from multimodal_emotion_detection_b200 import parallel as par
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 4, 8):
            r = [par.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            c = par.shard_counts(n, w)
            assert sum(c) == n and max(c) - min(c) <= 1
    with pytest.raises(ValueError):
        par.shard_range(10, 2, 2)


def test_split_format_round_trip(tmp_path):
    """The reference's on-disk split format (src/dataprocessing.py:450-490 writes it, src/data.py:51-84 reads it)."""
    rng = np.random.default_rng(0)
    def split(n):
        return {"audio": rng.normal(size=(n, 100, 1)).astype(np.float32),
                "video": rng.integers(0, 256, size=(n, 2, 16), dtype=np.uint8), "labels": rng.integers(0, 8, size=n)}
    tr, va, te = split(7), split(3), split(5)
    pkg.save_splits_to_disk(tr, va, te, str(tmp_path))
    for name, d in (("train", tr), ("val", va), ("test", te)):
        for k in ("audio", "video", "labels"):
            assert (np.load(tmp_path / name / f"{k}.npy") == d[k]).all()
    rd = pkg.SplitReader(str(tmp_path), ["audio", "video"], split="test", batch_size=2, modality_dropout=0.3, device="cpu")
    assert len(rd) == 3 and rd.modality_dropout == 0.0                    # dropout on the train split only (src/data.py:46)
    with pytest.raises(RuntimeError):
        next(iter(rd))                                                    # no CPU fallback
    with pytest.raises(FileNotFoundError):
        pkg.SplitReader(str(tmp_path), ["imu"], split="train")
    with pytest.raises(KeyError):
        pkg.save_splits_to_disk({"audio": tr["audio"], "labels": tr["labels"]}, va, te, str(tmp_path / "x"), ["audio", "imu"])
