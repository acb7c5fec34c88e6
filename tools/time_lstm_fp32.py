"""Time the fp32 LSTM encoder (the near-tie re-evaluation's kernels) at a flush-sized batch and dump the embedding so two
builds / settings can be compared bit for bit:  python tools/time_lstm_fp32.py B T out.npy"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import multimodal_emotion_detection_b200 as pkg

B, T, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
torch.manual_seed(0)
enc = pkg.SequenceEncoder(104, 256, 128, 2, "lstm").to("cuda:0").eval()
rng = np.random.default_rng(5)
x = torch.from_numpy(rng.normal(size=(B, T, 104)).astype(np.float32)).to("cuda:0")
lengths = torch.from_numpy(rng.integers(1, T + 1, size=B))
enc.precision = pkg._lib.PREC_FP32
y = enc(x, lengths)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    y = enc(x, lengths)
e1.record()
torch.cuda.synchronize()
print(f"B={B} T={T} MC={os.environ.get('AVHOT_LSTM_MC', '1')}: {e0.elapsed_time(e1) / 3:.2f} ms per forward")
np.save(out, y.cpu().numpy())
