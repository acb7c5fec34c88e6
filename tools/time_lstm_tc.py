"""Time the tensor-core LSTM encoder behind forward_from_waveform (front-end operand, DCT fold, fused layer kernels) for
1- and 2-layer stacks at the bench batch:  python tools/time_lstm_tc.py [B]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_emotion_detection_b200 as pkg

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128
dev = torch.device("cuda:0")
wave = pkg.synth.synth_audio(B, 0, 42, dev)
for layers in (1, 2):
    torch.manual_seed(0)
    enc = pkg.SequenceEncoder(104, 256, 128, layers, "lstm").to(dev).eval()
    enc.precision = pkg._lib.PREC_TC
    enc.forward_from_waveform(wave)
    torch.cuda.synchronize()
    pkg._lib.prof_reset()
    pkg._lib.prof_enable(True)
    for _ in range(3):
        enc.forward_from_waveform(wave)
    torch.cuda.synchronize()
    pkg._lib.prof_enable(False)
    rep = pkg._lib.prof_report()
    print(layers, "layer(s):", {k: round(v[0] / 3, 3) for k, v in rep.items()})
