#!/usr/bin/env python
"""ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) of
`python bench.py --no-e2e --no-cpu-baseline --no-extra --steps 2`  ->  the launches of one timed step of the main path
(profiles/<tag>/ncu_step_launches.csv) and profiles/ncu_traffic.json (DRAM bytes per step per bench.py stage name).
    python tools/summarize_ncu_launches.py gpurun_out/launches_r2.csv round2 [step_index]"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, tag = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 4            # 0-based index of the main-path step to report (3 warm-up + 2 timed)
STAGE = [("frontend_tc_kernel", "frontend"), ("lstm_fused_pair_kernel", "lstm_layer_fused_tc"), ("gemm_f16_ws_pair_kernel", "lstm_input_gemm_tc"), ("gemm_f16_ws_kernel", "lstm_input_gemm_tc"),
         ("lstm_rec_tc_kernel", "lstm_recurrence_tc"), ("gemm_tf32_tc_kernel", "frame_gemm"), ("gemm_cvt_f16_pair_kernel", "frame_gemm"), ("frame_tail_kernel", "frame_tail"),
         ("fusion_clip_kernel", "fusion_clip"), ("fusion_tc_kernel", "fusion_tc"), ("mc_reduce_kernel", "mc_reduce"),
         ("gemm_f32_kernel", "lstm_projection"), ("near_tie_flag_kernel", "near_tie_flag"), ("near_tie_begin_kernel", "near_tie_flag"),
         ("gather_flagged_kernel", "near_tie_gather")]
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
ix = {h: i for i, h in enumerate(rows[0])}
launches = collections.OrderedDict()
for r in rows[1:]:
    if not r[ix["ID"]].isdigit():
        continue
    L = launches.setdefault(int(r[ix["ID"]]), {"name": re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", ""),
                                               "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]]})
    L[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
ids = list(launches)
starts = [i for i in ids if "frontend_tc_kernel" in launches[i]["name"] and launches[i]["grid"].replace(" ", "").startswith("(148,")]
# a main-path step = from a full-grid front-end launch to the next mc_reduce
begin = starts[which]
step = []
for i in ids[ids.index(begin):]:
    step.append(i)
    if "mc_reduce" in launches[i]["name"]:
        break
out_csv = os.path.join(ROOT, "profiles", tag, "ncu_step_launches.csv")
tot = sum(launches[i]["gpu__time_duration.sum"] for i in step)
per_stage = collections.OrderedDict()
with open(out_csv, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["id", "kernel", "grid", "block", "time_us", "share_of_step", "dram_read_MB", "dram_write_MB", "bench_stage"])
    for i in step:
        L = launches[i]
        stage = next((s for k, s in STAGE if k in L["name"]), "")
        t, rd, wr = L["gpu__time_duration.sum"], L.get("dram__bytes_read.sum", 0.0), L.get("dram__bytes_write.sum", 0.0)
        w.writerow([i, L["name"], L["grid"], L["block"], f"{t / 1e3:.1f}", f"{t / tot:.4f}", f"{rd / 1e6:.1f}", f"{wr / 1e6:.1f}", stage])
        if stage:
            e = per_stage.setdefault(stage, {"time_us": 0.0, "bytes": 0.0, "launches": 0})
            e["time_us"] += t / 1e3
            e["bytes"] += rd + wr
            e["launches"] += 1
    w.writerow(["", "TOTAL (serialised, cold cache, unthrottled clocks: compare shares)", "", "", f"{tot / 1e3:.1f}", "1.0", "", "", ""])
traffic = {"batch": 18944, "precision": "tc", "source": f"profiles/{tag}/ncu_step_launches.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum, "
           "one timed step of python bench.py --no-e2e --no-cpu-baseline --no-extra --steps 2)",
           "bytes_per_step": {k: v["bytes"] for k, v in per_stage.items()},
           "ncu_time_us_per_step": {k: v["time_us"] for k, v in per_stage.items()},
           "launches_per_step": {k: v["launches"] for k, v in per_stage.items()}}
json.dump(traffic, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(open(out_csv).read())
print(json.dumps(traffic, indent=1))
