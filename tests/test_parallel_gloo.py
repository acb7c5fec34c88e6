"""N > 1 host logic on CPU: world_size-2 gloo processes shard a clip range and all-gather packed records."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_clips, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multimodal_emotion_detection_b200 import parallel as par
    lo, hi = par.shard_range(n_clips, rank, world)
    idx = torch.arange(lo, hi, dtype=torch.float32)
    packed = torch.stack([idx * 9 + c for c in range(9)], dim=1)       # record of clip g = [9g .. 9g+8]
    full = par.all_gather_packed(packed, n_clips)
    labels = par.labels_from_packed(full)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), full.numpy())
    assert labels.shape == (n_clips,)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [10, 11])
def test_all_gather_packed_world2(tmp_path, n_clips):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_clips, str(tmp_path)), nprocs=2, join=True)
    want = np.arange(n_clips * 9, dtype=np.float32).reshape(n_clips, 9)
    for r in range(2):
        got = np.load(tmp_path / f"r{r}.npy")
        assert got.shape == want.shape and (got == want).all()


def test_single_process_passthrough():
    from multimodal_emotion_detection_b200 import parallel as par
    x = torch.zeros(5, 9)
    assert par.all_gather_packed(x, 5) is x
    with pytest.raises(ValueError):
        par.all_gather_packed(x, 6)
