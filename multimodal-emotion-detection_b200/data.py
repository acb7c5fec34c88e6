"""The reference's on-disk split format, either side of the hot path.

save_splits_to_disk  src/dataprocessing.py:450-490: <out_root>/{train,val,test}/{<modality>.npy, labels.npy}.
SplitReader          what MultimodalDataset + collate_multimodal + the DataLoader hand to the model
                     (src/data.py:19-122, :191-211): batches (features: Dict[str, Tensor], labels, mask) - here read from
                     memory-mapped .npy files through pinned staging buffers, copied to the device on a side stream
                     one batch ahead, with the modality-dropout mask (src/data.py:108-118, train split only) drawn on
                     the device by the Philox spec (avhot_modality_mask, keyed by the sample's index in the split).
                     uint8 video files stay uint8 (the frame encoder divides by 255 on the device).
The batches are what AVHotPath / MCDropoutUncertainty / compute_calibration_metrics take.
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np
import torch

from . import _rng, synth


def save_splits_to_disk(train_data: Dict[str, np.ndarray], val_data: Dict[str, np.ndarray],
                        test_data: Dict[str, np.ndarray], out_root: str, modalities: Optional[List[str]] = None) -> None:
    out_root = Path(out_root)
    out_root.mkdir(parents=True, exist_ok=True)
    if modalities is None:
        modalities = [k for k in train_data.keys() if k != "labels"]                    # :470-471
    for split_name, data_dict in (("train", train_data), ("val", val_data), ("test", test_data)):
        split_dir = out_root / split_name
        split_dir.mkdir(parents=True, exist_ok=True)
        for m in modalities:
            if m not in data_dict:
                raise KeyError(f"Modality '{m}' missing from {split_name} data.")        # :479
            np.save(split_dir / f"{m}.npy", data_dict[m])
        np.save(split_dir / "labels.npy", data_dict["labels"])


class SplitReader:
    """Iterable of (features, labels, mask) device batches over one split, in file order."""

    def __init__(self, data_dir: str, modalities: List[str], split: str = "train", batch_size: int = 32,
                 modality_dropout: float = 0.0, device="cuda", seed: int = _rng.DEFAULT_SEED):
        self.modalities = list(modalities)
        self.split = split
        self.batch_size = int(batch_size)
        self.modality_dropout = modality_dropout if split == "train" else 0.0         # src/data.py:46
        self.device = torch.device(device)
        self.seed = seed
        split_dir = Path(data_dir) / split
        self.data = {}
        for m in self.modalities:
            f = split_dir / f"{m}.npy"
            if not f.exists():
                raise FileNotFoundError(f"Modality file not found: {f}")                # :74
            self.data[m] = np.load(f, mmap_mode="r")
        lf = split_dir / "labels.npy"
        if not lf.exists():
            raise FileNotFoundError(f"Labels file not found: {lf}")                     # :81
        self.labels = np.load(lf)
        for m, a in self.data.items():
            if len(a) != len(self.labels):
                raise ValueError(f"{m}.npy has {len(a)} samples, labels.npy {len(self.labels)}")

    def __len__(self) -> int:
        return (len(self.labels) + self.batch_size - 1) // self.batch_size

    def _stage(self, lo: int, hi: int, bufs) -> Tuple[Dict[str, torch.Tensor], torch.Tensor]:
        out = {}
        for m, a in self.data.items():
            keep_u8 = a.dtype == np.uint8
            buf = bufs[m][:hi - lo]
            np.copyto(buf.numpy(), a[lo:hi] if keep_u8 else np.asarray(a[lo:hi], dtype=np.float32))   # .float(), :100
            out[m] = buf
        lab = bufs["labels"][:hi - lo]
        np.copyto(lab.numpy(), self.labels[lo:hi].astype(np.int64))
        return out, lab

    def __iter__(self) -> Iterator[Tuple[Dict[str, torch.Tensor], torch.Tensor, torch.Tensor]]:
        n, bs, dev = len(self.labels), self.batch_size, self.device
        if dev.type != "cuda":
            raise RuntimeError("SplitReader feeds the CUDA path; there is no CPU fallback")
        sets = []
        for _ in range(2):                                       # two pinned staging sets, alternated
            b = {m: torch.empty((bs,) + tuple(a.shape[1:]), dtype=torch.uint8 if a.dtype == np.uint8 else torch.float32,
                                pin_memory=True) for m, a in self.data.items()}
            b["labels"] = torch.empty(bs, dtype=torch.int64, pin_memory=True)
            sets.append(b)
        copy_stream = torch.cuda.Stream(device=dev)
        done = [None, None]                                       # copy-complete events per staging set

        def upload(i: int):
            lo, hi = i * bs, min(n, (i + 1) * bs)
            s = i % 2
            if done[s] is not None:
                done[s].synchronize()                             # the staging set is free again
            host, lab = self._stage(lo, hi, sets[s])
            with torch.cuda.stream(copy_stream):
                feats = {m: t.to(dev, non_blocking=True) for m, t in host.items()}
                labels = lab.to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            done[s] = ev
            return feats, labels, ev, lo, hi

        nb = len(self)
        nxt = upload(0) if nb else None
        for i in range(nb):
            feats, labels, ev, lo, hi = nxt
            nxt = upload(i + 1) if i + 1 < nb else None           # one batch ahead
            torch.cuda.current_stream(dev).wait_event(ev)
            for t in list(feats.values()) + [labels]:
                t.record_stream(torch.cuda.current_stream(dev))
            if self.modality_dropout > 0:
                mask = synth.modality_mask(hi - lo, lo, self.modality_dropout, self.seed, dev, len(self.modalities))
            else:
                mask = torch.ones(hi - lo, len(self.modalities), device=dev)             # :105
            yield feats, labels, mask
