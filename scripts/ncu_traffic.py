#!/usr/bin/env python
"""DRAM bytes per launch of every kernel of one bench step, from an `ncu --set full` report of bench.py
(scripts/gpu_profiles.sh) -> JSON {bench kernel name: {"dram_bytes": read+write, "tensor_pipe_active_pct": ...}}.

    python scripts/ncu_traffic.py gpurun_out/r01c_ncu_all.ncu-rep > profiles/r01c_traffic.json
"""
import csv
import io
import json
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}
tmult = {"us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}
names = [("stft_hop1", "stft_hop1"), ("if_reassign", "if_reassign"), ("stats_finalize", "stats_finalize"), ("normalise", "normalise"),
         ("split_planes", "split_planes"), ("tc_inproj", "tc_inproj"), ("tc_recurrent", "tc_recurrent"), ("head_kernel", "head"),
         ("confusion_kernel", "confusion"), ("metrics_kernel", "metrics")]
out, n_inproj = {}, 0
for r in rows[2:]:
    kname = r[idx["Kernel Name"]]
    key = next((v for k, v in names if k in kname), None)
    if key is None:
        continue
    if key == "tc_inproj":          # <3, 8, ..> = the layer-1 projection (stand-in / HSSB_FUSE_X=0 path), <4, 4, ..> = layer 2's
        key = "tc_inproj_l0" if "<3, 8" in kname or "(int)3, (int)8" in kname else "tc_inproj_l1"
        n_inproj += 1
    if key == "tc_recurrent":       # <S, publish, EW, fused, train>: fused = layer 1 (input projection in the kernel), else layer 2
        args = kname[kname.index("<") + 1:kname.index(">")].replace("(int)", "").replace("(bool)", "").split(",")
        key = "tc_recurrent_l1" if args[3].strip() in ("1", "true") else "tc_recurrent_l2"
    if key.startswith("tc_") and float(r[idx["gpu__time_duration.sum"]]) * tmult[units[idx["gpu__time_duration.sum"]]] < 0.02:
        continue                    # a stand-in launch of the input-range guard that exited at once
    b = float(r[idx["dram__bytes_read.sum"]]) * mult[units[idx["dram__bytes_read.sum"]]] + \
        float(r[idx["dram__bytes_write.sum"]]) * mult[units[idx["dram__bytes_write.sum"]]]
    e = out.setdefault(key, {"launches": 0, "dram_bytes": 0.0, "ncu_ms": 0.0, "tensor_pipe_active_pct": 0.0})
    e["launches"] += 1
    e["dram_bytes"] += b
    e["ncu_ms"] += float(r[idx["gpu__time_duration.sum"]]) * tmult[units[idx["gpu__time_duration.sum"]]]
    e["tensor_pipe_active_pct"] += float(r[idx["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])
for e in out.values():
    n = e["launches"]
    e["dram_bytes_per_launch"] = e.pop("dram_bytes") / n
    e["ncu_ms_per_launch"] = e.pop("ncu_ms") / n
    e["tensor_pipe_active_pct"] /= n
print(json.dumps({"source": sys.argv[1].split("/")[-1], "workload": "bench.py default: 512 windows x 2000 samples, 1 GPU", "kernels": out}, indent=1))
