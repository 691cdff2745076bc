#!/usr/bin/env python
"""profiles/<tag>_sanitizer.txt from the logs scripts/sanitize_r2.sh leaves under gpurun_out/r2/."""
import glob
import os
import re
import sys

rows = []
for path in sorted(glob.glob("gpurun_out/r2/sanitize_*_*.txt")):
    tool, what = re.match(r"sanitize_([a-z]+)_(.+)\.txt", os.path.basename(path)).groups()
    text = open(path, errors="replace").read()
    summary = re.findall(r"=========\s*((?:ERROR|RACECHECK) SUMMARY:[^\n]*)", text)
    done = [ln.strip() for ln in text.splitlines() if re.match(r"^(fsst|lstm|overlap|pipeline|train) ", ln)]
    rows.append((tool, what, summary[-1] if summary else "no summary (killed by the time limit)", "; ".join(done)))
order = {"memcheck": 0, "synccheck": 1, "racecheck": 2}
rows.sort(key=lambda r: (order.get(r[0], 9), r[1]))
print("# compute-sanitizer passes (scripts/sanitize_r2.sh -> scripts/sanitize_r2.py workloads on a B200), summarised by scripts/sanitize_summary.py")
print("# workloads: fsst (3x300, 37x170 windows), lstm (70x24, 200x12: several clusters, ragged groups), overlap (40 x 1030: the projection")
print("#            as launches M / A / B around the recurrences), pipeline (the same with the next batch's FSST behind the side gate), train (tensor-core training step: K4 + K5m TRAIN forward, K5b backward,")
print("#            TF32 splits, fused head + loss, clip + Adam; shapes in the result column), train_simt (the fp32 cluster kernels)")
for tool, what, summary, done in rows:
    print(f"{tool:10s} {what:11s} {summary}   [{done}]")
