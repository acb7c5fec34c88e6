"""Debug driver: the LSTM encoder on the tensor-core path (fused top layer) against the fp32 kernels."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import multimodal_emotion_detection_b200 as pkg

B = int(sys.argv[1]) if len(sys.argv) > 1 else 300
T = int(sys.argv[2]) if len(sys.argv) > 2 else 20
L = int(sys.argv[3]) if len(sys.argv) > 3 else 2
torch.manual_seed(0)
enc = pkg.SequenceEncoder(104, 256, 128, L, "lstm").to("cuda:0").eval()
rng = np.random.default_rng(5)
x = torch.from_numpy(rng.normal(size=(B, T, 104)).astype(np.float32)).to("cuda:0")
lengths = torch.from_numpy(rng.integers(1, T + 1, size=B))
enc.precision = pkg._lib.PREC_TC
out = enc(x, lengths)
torch.cuda.synchronize()
enc.precision = pkg._lib.PREC_FP32
ref = enc(x, lengths)
torch.cuda.synchronize()
err = (out - ref).abs().max().item() / ref.abs().max().item()
print("B", B, "T", T, "L", L, "rel err tc vs fp32:", err)
