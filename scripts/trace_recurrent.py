#!/usr/bin/env python
"""Per-role clock64 trace of the tcgen05 recurrence kernel (diagnostic; see hssb_debug_trace).

    python scripts/trace_recurrent.py [B] [T] [steps]

Needs a diagnostic build (the stamps are compiled out of the default one): HSSB_TRACE=1 python __graft_entry__.py --force
Runs one eval forward with tracing on and prints, per sub-tile, the mean cycle offsets of every
event relative to the MMA thread's "h_full seen" stamp of the same step, plus the step period.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heart-sounds-segmentation_b200"))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from hss import _lib
from hss.model.segmenter import HeartSoundSegmenter

EV = ["mma_hfull", "mma_issued", "epi_dfull", "epi_act", "epi_cell", "epi_image", "epi_copies"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T = int(sys.argv[2]) if len(sys.argv) > 2 else 256
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 64
torch.manual_seed(0)
m = HeartSoundSegmenter(input_size=44, batch_size=B).eval()
x = torch.randn(B, T, 44, device="cuda")
m.predict(x)
buf = torch.zeros(steps * 4 * 16, dtype=torch.int64, device="cuda")
lib = _lib.lib()
print("max co-resident clusters (32x2):", lib.hssb_debug_max_clusters())
lib.hssb_debug_trace(buf.data_ptr(), steps)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
_lib.prof_enable(True); _lib.prof_read()
e0.record(); m.predict(x); e1.record()
torch.cuda.synchronize()
print("geom", os.environ.get("HSSB_RC_GEOM", "auto"), "B", B, "T", T, "forward ms", e0.elapsed_time(e1), {k: round(v[1], 3) for k, v in _lib.prof_read().items()})
lib.hssb_debug_trace(None, 0)
tr = buf.cpu().numpy().reshape(steps, 4, 16).astype(np.float64)   # last launch = layer 2
for s in range(4):
    if tr[:, s, 0].max() == 0:
        continue
    base = tr[:, s, 0]
    period = np.diff(base)[8:]
    if period.size == 0:
        continue
    print(f"sub-tile {s}: step period mean {period.mean():.0f} cyc (min {period.min():.0f}, max {period.max():.0f})")
    for e, name in enumerate(EV):
        d = (tr[8:, s, e] - base[8:])
        print(f"   {name:16s} {d.mean():9.0f}  (min {d.min():9.0f} max {d.max():9.0f})")
# absolute timeline of a few steps (all sub-tiles interleaved), cycles relative to step 20's first event
if os.environ.get("HSSB_TRACE_TIMELINE"):
    ev = []
    for st in range(20, 23):
        for s in range(4):
            for e, name in enumerate(EV):
                if tr[st, s, e] > 0:
                    ev.append((tr[st, s, e], f"s{s} t{st} {name}"))
    ev.sort()
    t0 = ev[0][0]
    for c, n in ev:
        print(f"{c - t0:8.0f}  {n}")
