"""tests/golden/seq_golden.npz (made by tests/golden/make_golden_seq.py from the reference's own classes): loaders."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAX_LEN = 4096


def load():
    return np.load(os.path.join(ROOT, "tests", "golden", "seq_golden.npz"))


def state_dict(g, prefix, keep_prefix=False):
    """Weights stored as "w.<prefix><state-dict key>"; the transformer's pos_embedding is stored truncated (the rest
    is zeros)."""
    sd = {}
    for k in g.files:
        if not k.startswith("w." + prefix):
            continue
        name = k[2 + len(prefix):]
        v = torch.from_numpy(g[k])
        if name == "pos_embedding.weight":
            full = torch.zeros(MAX_LEN, v.shape[1])
            full[:v.shape[0]] = v
            v = full
        sd[prefix + name if keep_prefix else name] = v
    return sd


PAIR_DIMS = {"audio": 128, "video": 128, "imu": 64}
