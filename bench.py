#!/usr/bin/env python
"""Benchmark of the per-clip audio-visual inference hot path (BASELINE.json metric: clips/sec, front-end + fused
forward + T-pass MC-dropout uncertainty).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] --steps K --warmup W   # the reference's CPU path (oracle port)

A step = one batch of synthetic RAVDESS-shaped clips per GPU through: log-mel/MFCC front-end -> LSTM audio encoder
+ frame encoder -> hybrid cross-modal attention fusion -> T = 32 MC-dropout passes -> (mean logits, uncertainty,
label) per clip; with N > 1 ranks (one process per GPU, weak scaling) the packed (B, 9) records are all-gathered
over NCCL.  `value` times the steps with inputs resident in HBM; `e2e` times the same call with HOST (pinned)
buffers, host->device copies and the device->host read of the result inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 42
N_SAMPLES, N_FRAMES, FRAME_DIM, N_CLASSES = 48000, 16, 112 * 112, 8
T_FEATS = 1 + N_SAMPLES // 160
P_MODALITY = 0.05
# algorithmic work per clip (derivations in DESIGN.md)
FLOPS_LSTM_RECURRENT_PER_LAYER = T_FEATS * 2 * 1024 * 256          # h_{t-1} @ W_hh^T, one layer
FLOPS_LSTM = 537_542_656
FLOPS_FRAME_GEMM = 16 * 2 * FRAME_DIM * 256
BYTES_FRONTEND = N_SAMPLES * 4 + T_FEATS * 104 * 4
BYTES_FRAMES = N_FRAMES * FRAME_DIM * 4
BYTES_IN_PER_CLIP = N_SAMPLES * 4 + BYTES_FRAMES + 2 * 4
BYTES_OUT_PER_CLIP = (N_CLASSES + 1) * 4


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's execution of the path (oracle port), on the host cores
# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_step(n_clips: int, passes: int, clip0: int, sd, threads: int):
    """One bounded sample: n_clips synthetic clips through the front-end restatement (once: the reference would
    precompute features offline) and MCDropoutUncertainty as the reference executes it (T full forwards,
    torch.nn.LSTM).  Returns (seconds, mean_logits)."""
    import numpy as np
    import torch
    from oracle import frontend as ofe, model as om, philox as px, synth as osynth
    torch.set_num_threads(threads)
    ids = np.arange(n_clips, dtype=np.uint64) + np.uint64(clip0)
    wave = osynth.synth_audio(ids, SEED)                       # input generation is not timed
    frames = torch.from_numpy(osynth.synth_video(ids, SEED))
    mask = torch.from_numpy(px.modality_mask(ids, 2, P_MODALITY, SEED))
    t0 = time.perf_counter()
    feats = torch.from_numpy(ofe.logmel_mfcc(wave))
    mean, unc, _ = om.mc_dropout_forward_as_reference(feats, frames, mask, sd, passes, SEED, ids, "hybrid")
    return time.perf_counter() - t0, mean


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from oracle import model as om, synth as osynth
    threads = os.cpu_count() or 1
    sd = om.to_torch(osynth.synth_weights(SEED))
    n = args.cpu_clips
    for w in range(args.warmup):
        cpu_reference_step(n, args.passes, w * n, sd, threads)
    total = 0.0
    for s in range(args.steps):
        dt, _ = cpu_reference_step(n, args.passes, (args.warmup + s) * n, sd, threads)
        total += dt
    value = n * args.steps / total
    sample = (f"{n} synthetic clips per step: fp64 numpy front-end restatement once + MCDropoutUncertainty as the "
              f"reference runs it (T={args.passes} full forwards, torch.nn.LSTM + torch CPU ops, fp32, "
              f"{threads} threads, torch {torch.__version__})")
    line = {"impl": "reference", "metric": "clips_per_sec", "value": value, "unit": "clips/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, n),
            "cpu_baseline": {"value": value, "unit": "clips/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def bind_to_gpu_numa_node(torch, index):
    """One process per GPU: run this process (and so allocate its pinned host buffers, first touch) on the CPU socket the
    GPU hangs off, so the host->device copies of the e2e leg do not cross the socket interconnect.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(index).uuid)
        handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        bus_id = pynvml.nvmlDeviceGetPciInfo(handle).busId
        bus_id = (bus_id.decode() if isinstance(bus_id, bytes) else bus_id).lower()
        if len(bus_id.split(":")[0]) == 8:               # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus_id = bus_id[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus_id}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:                                  # noqa: BLE001 - placement is an optimisation, never a failure
        pass
    return None


def workload_config(args, batch):
    return {"workload": "BASELINE.json configs[4] shard: synthetic RAVDESS-shaped clips (3 s 16 kHz audio, 16 frames "
                        "112x112 fp32) -> log-mel/MFCC front-end (n_fft 512, hop 160, 64 mel + 40 MFCC) -> LSTM 2x256 "
                        "+ frame encoder -> hybrid cross-modal attention fusion -> MC-dropout uncertainty",
            "clips_per_gpu_per_step": batch, "mc_passes": args.passes, "modality_dropout_p": P_MODALITY,
            "classes": N_CLASSES, "precision": args.precision, "parallelism": f"clip-sharded x{args.gpus}, all-gather of (B,9) fp32",
            "l2_policy": f"inputs larger than L2 ({batch * BYTES_IN_PER_CLIP / 1e9:.1f} GB per step per GPU), two "
                         "batches alternate"}


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def run_b200_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun --nproc-per-node {args.gpus}")
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    numa = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    if world > 1:
        # stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION) off it
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    import multimodal_emotion_detection_b200 as pkg
    from multimodal_emotion_detection_b200 import parallel as par

    pkg._lib.check_device(dev)
    B, T = args.batch, args.passes
    # random-init weights of the architecture: torch's default init under the reference's seed (configs/base.yaml:4,
    # src/train.py:472).  The product arm touches nothing under oracle/; the CPU leg below is handed these tensors.
    torch.manual_seed(SEED)
    model = pkg.AVHotPath("hybrid")
    sd_cpu = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.to(dev).eval().set_precision(args.precision)
    model.fold_video_projection = not args.no_fold
    pkg.MCDropoutUncertainty._set_dropout_train(model)            # what MCDropoutUncertainty.forward does (:58-59)

    def clip0_of(step):                                            # global clip index of this rank's batch
        return (step * world + rank) * B

    # two resident batches, alternated (each 4 GB >> 126 MB of L2)
    pool = []
    for j in range(2):
        c0 = clip0_of(j)
        pool.append(({"audio": pkg.synth.synth_audio(B, c0, SEED, dev), "video": pkg.synth.synth_video(B, c0, SEED, dev)},
                     pkg.synth.modality_mask(B, c0, P_MODALITY, SEED, dev), c0))
    torch.cuda.synchronize()

    def step(j):
        feats, mask, c0 = pool[j % 2]
        mean, unc, labels, packed = model.mc_forward(feats, mask, num_samples=T, seed=SEED, clip0=c0)
        if world > 1:
            packed = par.all_gather_packed(packed, B * world)
        return packed

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for j in range(args.warmup):
        step(j)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    pkg._lib.prof_reset()
    pkg._lib.prof_enable(True)
    launches0 = pkg._lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for j in range(args.steps):
        out = step(args.warmup + j)
    e1.record()
    barrier()
    launches = pkg._lib.launch_count() - launches0
    pkg._lib.prof_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    value = world * B * args.steps / (ms / 1e3)
    prof = pkg._lib.prof_report()
    pkg._lib.prof_reset()

    # ---- end to end through the public API with host buffers --------------------------------------------------
    e2e = None if args.no_e2e else run_e2e(args, pkg, par, model, dev, world, rank, B, T, clip0_of, barrier)
    # the same leg with the frames kept as the uint8 pixels they are decoded to (src/dataprocessing.py:261-265
    # divides by 255 on the host; here the division happens on the device): a quarter of the dominant H2D stream
    e2e_u8 = None if args.no_e2e else run_e2e(args, pkg, par, model, dev, world, rank, B, T, clip0_of, barrier, u8_frames=True)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = load_peaks()
    stages = {}
    for name, (tot, n) in sorted(prof.items()):
        stages[name] = {"ms_per_step": tot / args.steps, "launches_per_step": n / args.steps}
    # ---- rooflines: algorithmic bytes / FLOPs per launch over the CUDA-event time of the launch
    tc = args.precision == "tc"
    pitch = 320                                   # fp16 halves per row of the split LSTM operand (3 * 104 + 2, padded)
    Bp = (B + 127) // 128 * 128
    xp_bytes = Bp * T_FEATS * 4096                # one layer's input projections, fp32
    algo = {   # name -> (bound, algorithmic bytes per step over all its launches, algorithmic FLOPs per step)
        "frontend": ("hbm", B * N_SAMPLES * 4 + (Bp * T_FEATS * pitch * 2 if tc else B * T_FEATS * 104 * 4), 0),
        "frame_gemm": ("hbm", B * BYTES_FRAMES + B * 16 * 256 * 4, B * FLOPS_FRAME_GEMM),
        # xp is fp16 for both layers -> 2 x half an fp32 layer, written by the GEMM and read by the recurrence
        "lstm_input_gemm_tc": ("hbm", xp_bytes + Bp * T_FEATS * (pitch + 256) * 2, B * (FLOPS_LSTM - 2 * FLOPS_LSTM_RECURRENT_PER_LAYER)),
        "lstm_recurrence_tc": ("hbm", xp_bytes + Bp * T_FEATS * 256 * 2, 2 * B * FLOPS_LSTM_RECURRENT_PER_LAYER),
        "lstm_recurrence": ("tensor", 0, 2 * B * FLOPS_LSTM_RECURRENT_PER_LAYER),
        "lstm_input_gemm": ("tensor", 0, B * (FLOPS_LSTM - 2 * FLOPS_LSTM_RECURRENT_PER_LAYER)),
    }
    extra = {}
    for name, (bound, nbytes, nflops) in algo.items():
        tot, n = prof.get(name, (0.0, 0))
        if not n:
            continue
        sec = tot / args.steps / 1e3
        ent = {"bound": bound, "ms_per_step": tot / args.steps, "launches_per_step": n / args.steps}
        if nbytes:
            ent.update({"hbm_gbs": nbytes / sec / 1e9, "hbm_frac": nbytes / sec / 1e9 / peaks["hbm_gbs"],
                        "algorithmic_bytes_per_step": nbytes})
        if nflops:
            ent.update({"tflops": nflops / sec / 1e12, "tensor_frac": nflops / sec / 1e12 / peaks["tf_sustained"]})
        if nbytes and nflops:       # a kernel that streams and contracts: the roofline it sits closer to is its bound
            ent["bound"] = "tensor" if ent["tensor_frac"] > ent["hbm_frac"] else "hbm"
        extra[name] = ent
    # the dominant kernel = the one with the most device time per step among the stages with a stated roofline
    roofline = None
    if extra:
        dom = max(extra, key=lambda k: extra[k]["ms_per_step"])
        e = extra[dom]
        tot, n = prof[dom]
        if e["bound"] == "hbm":
            roofline = {"kernel": dom, "bound": "hbm", "achieved": e["hbm_gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": e["hbm_frac"], "traffic": None}
        else:
            roofline = {"kernel": dom, "bound": "tensor", "achieved": e["tflops"], "peak": peaks["tf_sustained"],
                        "unit": "TFLOP/s", "frac": e["tensor_frac"], "traffic": None}
        # DRAM traffic of that kernel per launch from the committed ncu capture of the same workload, if there is one
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                tr = json.load(f)
            if tr.get("batch") == B and tr.get("precision") == args.precision:
                roofline["traffic"] = tr["bytes_per_launch"].get(dom)
                roofline["traffic_source"] = "profiles/ncu_traffic.json"
                roofline["algorithmic_bytes_per_launch"] = e.get("algorithmic_bytes_per_step", 0) / (n / args.steps)
        except (OSError, ValueError):
            pass
        roofline.update({"peak_source": peaks["source"], "avg_launch_ms": tot / n,
                         "share_of_step": tot / args.steps / (ms / args.steps),
                         "note": "dominant kernel by device time per step; every stage is in stage_rooflines"})

    # labels of the timed precision against the fp32 CUDA path (itself checked against the oracle) on one batch
    parity = None
    if args.precision != "fp32":
        feats, mask, c0 = pool[0]
        nchk = min(B, 1024)
        sub = {k: v[:nchk] for k, v in feats.items()}
        a = model.mc_forward(sub, mask[:nchk], num_samples=T, seed=SEED, clip0=c0)
        model.set_precision("fp32")
        b = model.mc_forward(sub, mask[:nchk], num_samples=T, seed=SEED, clip0=c0)
        model.set_precision(args.precision)
        parity = {"clips": nchk, "label_mismatches_vs_fp32_path": int((a[2] != b[2]).sum().item()),
                  "mean_logits_rel_err_vs_fp32_path": float((a[0] - b[0]).abs().max() / b[0].abs().max()),
                  "uncertainty_rel_err_vs_fp32_path": float((a[1] - b[1]).abs().max() / b[1].abs().max())}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import model as om
        sd = sd_cpu                                                  # AVHotPath("hybrid") keys = the oracle's (Appendix A)
        threads = os.cpu_count() or 1
        cpu_reference_step(min(8, args.cpu_clips), 2, 0, sd, threads)          # warm-up
        dt, cpu_mean = cpu_reference_step(args.cpu_clips, T, clip0_of(0), sd, threads)
        # same clips through the CUDA path: labels must agree
        feats, mask, c0 = pool[0]
        n = args.cpu_clips
        g_mean = model.mc_forward({k: v[:n] for k, v in feats.items()}, mask[:n], num_samples=T, seed=SEED, clip0=c0)[0]
        rel = float((g_mean.cpu() - cpu_mean).abs().max() / cpu_mean.abs().max())
        cpu_baseline = {"value": args.cpu_clips / dt, "unit": "clips/s", "cores": threads, "kind": "port",
                        "sample": f"{args.cpu_clips} clips of the same workload, T={T} full forwards as "
                                  f"src/uncertainty.py:62-65 runs them (torch.nn.LSTM + torch CPU ops fp32, numpy fp64 "
                                  f"front-end once), {dt:.1f} s",
                        "labels_equal_to_gpu": bool((g_mean.argmax(1).cpu() == cpu_mean.argmax(1)).all()),
                        "mean_logits_rel_err_vs_gpu": rel}

    line = {"metric": "clips_per_sec", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "tf32/f16 tensor-core operands, f32 accumulate",
            "data": "synthetic", "config": workload_config(args, B),
            "e2e": e2e, "e2e_u8_frames": e2e_u8, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "parity": parity, "stages": stages, "stage_rooflines": extra,
            "host_numa_node_rank0": numa}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_e2e(args, pkg, par, model, dev, world, rank, B, T, clip0_of, barrier, u8_frames=False):
    """Same step through the public API (AVHotPath.mc_forward = MCDropoutUncertainty's fast path) with inputs in
    pinned HOST memory: every step uploads its B clips host->device in sub-batches of --e2e-chunk clips and reads the
    packed results back; the upload of sub-batch i+1 overlaps the kernels of sub-batch i on a second stream (two
    device staging slots).  The pinned source ring holds two sub-batches (real copies every time; clip indices for
    the Philox masks advance normally)."""
    import torch
    import torch.distributed as dist
    C = min(args.e2e_chunk, B)
    n_sub = (B + C - 1) // C
    host = []
    for j in range(2):
        c0 = clip0_of(j)
        a = pkg.synth.synth_audio(C, c0, SEED, dev)
        v = pkg.synth.synth_video(C, c0, SEED, dev, dtype=torch.uint8 if u8_frames else torch.float32)
        m = pkg.synth.modality_mask(C, c0, P_MODALITY, SEED, dev)
        ha = torch.empty(a.shape, dtype=a.dtype, pin_memory=True).copy_(a)
        hv = torch.empty(v.shape, dtype=v.dtype, pin_memory=True).copy_(v)
        hm = torch.empty(m.shape, dtype=m.dtype, pin_memory=True).copy_(m)
        host.append((ha, hv, hm))
        del a, v, m
    slots = [(torch.empty(C, N_SAMPLES, device=dev),
              torch.empty(C, N_FRAMES, FRAME_DIM, device=dev, dtype=torch.uint8 if u8_frames else torch.float32),
              torch.empty(C, 2, device=dev)) for _ in range(2)]
    packed_dev = torch.empty(B, N_CLASSES + 1, device=dev)
    result_host = torch.empty(B * world, N_CLASSES + 1, pin_memory=True)
    copy_stream = torch.cuda.Stream(dev)
    main = torch.cuda.current_stream(dev)
    uploaded = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    counter = [0]

    def run(n_steps, step0):
        for j in range(n_steps):
            base = clip0_of(step0 + j)
            for i in range(n_sub):
                n = min(C, B - i * C)
                k = counter[0]
                counter[0] += 1
                s = k % 2
                ha, hv, hm = host[s]
                with torch.cuda.stream(copy_stream):
                    if k >= 2:
                        copy_stream.wait_event(consumed[s])
                    slots[s][0][:n].copy_(ha[:n], non_blocking=True)
                    slots[s][1][:n].copy_(hv[:n], non_blocking=True)
                    slots[s][2][:n].copy_(hm[:n], non_blocking=True)
                    uploaded[s].record(copy_stream)
                main.wait_event(uploaded[s])
                packed = model.mc_forward({"audio": slots[s][0][:n], "video": slots[s][1][:n]}, slots[s][2][:n],
                                          num_samples=T, seed=SEED, clip0=base + i * C)[3]
                consumed[s].record(main)
                packed_dev[i * C:i * C + n].copy_(packed)
            out = par.all_gather_packed(packed_dev, B * world) if world > 1 else packed_dev
            result_host.copy_(out, non_blocking=True)

    run(1, 0)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    copy_stream.wait_event(e0)
    run(args.steps, 1)
    e1.record(main)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    in_bytes = N_SAMPLES * 4 + N_FRAMES * FRAME_DIM * (1 if u8_frames else 4) + 2 * 4
    return {"value": world * B * args.steps / (ms / 1e3), "unit": "clips/s", "ms_per_step": ms / args.steps,
            "h2d_bytes_per_step": B * in_bytes, "d2h_bytes_per_step": B * world * BYTES_OUT_PER_CLIP,
            "api": f"AVHotPath.mc_forward (fast path of MCDropoutUncertainty.forward), pinned host buffers "
                   f"(fp32 audio, {'uint8' if u8_frames else 'fp32'} frames), {n_sub} sub-batches of <= {C} clips per "
                   f"step, uploads double-buffered on a copy stream"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=148 * 128,
                    help="clips per GPU per step (default: one 128-clip CTA of the recurrent kernel per SM)")
    ap.add_argument("--e2e-chunk", type=int, default=2048, help="clips per host->device sub-batch in the e2e leg")
    ap.add_argument("--passes", type=int, default=32, help="MC-dropout passes T")
    ap.add_argument("--cpu-clips", type=int, default=384, help="clips in the bounded CPU sample (~10 s of host work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer legs (kernel A/B runs)")
    ap.add_argument("--no-fold", action="store_true",
                    help="keep z_v materialised instead of folding FrameEncoder.projection[1] into the fusion GEMM (A/B)")
    ap.add_argument("--precision", default="tc", choices=["tc", "fp32"],
                    help="tc: large contractions on tcgen05 (TF32 / fp16 operands); fp32: CUDA-core fp32 everywhere")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3                                         # timing rule: W >= 3
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
