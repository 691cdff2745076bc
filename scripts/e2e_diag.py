#!/usr/bin/env python
"""Diagnostic: where does the end-to-end (pinned host in, labels out) step spend its time?

    python scripts/e2e_diag.py [windows] [steps]

Prints, per step, wall-clock and CUDA-event durations of the H2D copy, the FSST, the BiLSTM and the
D2H copies, and the per-kernel profile of the e2e loop next to that of the device-resident loop.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heart-sounds-segmentation_b200"))
sys.path.insert(0, ROOT)
import torch
from hss import _lib
from hss.model.segmenter import HeartSoundSegmenter
from hss.sharding import confusion_counts
from hss.transforms import FSST
from workloads import reference_window, synthetic_targets, tiled_windows

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
N = 2000
dev = torch.device("cuda", 0)
x_host = torch.from_numpy(tiled_windows(B, N, 1000.0, 68)).pin_memory()
y_dev = torch.from_numpy(synthetic_targets(B, N)).to(dev)
fsst = FSST(1000.0, window=reference_window(128), truncate_freq=(25, 200), stack=True)
torch.manual_seed(68)
model = HeartSoundSegmenter(input_size=44, batch_size=B).eval()
labels_host = torch.empty((B, N), dtype=torch.int32).pin_memory()


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


x_dev = x_host.to(dev)
for _ in range(3):
    labels = model.predict(fsst.batch(x_dev))
torch.cuda.synchronize()

PROF = os.environ.get("DIAG_PROF", "1") == "1"
for mode in ("device", "e2e", "e2e-nosync"):
    _lib.prof_enable(PROF)
    _lib.prof_read()
    t_all0 = time.perf_counter()
    for s in range(steps):
        w0 = time.perf_counter()
        e0 = ev()
        x = x_host.to(dev, non_blocking=True) if mode != "device" else x_dev
        e1 = ev()
        feats = fsst.batch(x)
        e2 = ev()
        w1 = time.perf_counter()
        labels = model.predict(feats)
        e3 = ev()
        w2 = time.perf_counter()
        cm = confusion_counts(labels, y_dev)
        if mode != "device":
            labels_host.copy_(labels, non_blocking=True)
            if mode == "e2e":
                cm_host = cm.cpu()
        e4 = ev()
        w3 = time.perf_counter()
        torch.cuda.synchronize()
        w4 = time.perf_counter()
        print(f"{mode} step {s}: events h2d {e0.elapsed_time(e1):.3f} fsst {e1.elapsed_time(e2):.3f} lstm {e2.elapsed_time(e3):.3f} "
              f"out {e3.elapsed_time(e4):.3f} total {e0.elapsed_time(e4):.3f} ms | wall launch-fsst {1e3 * (w1 - w0):.3f} "
              f"launch-lstm {1e3 * (w2 - w1):.3f} out {1e3 * (w3 - w2):.3f} sync {1e3 * (w4 - w3):.3f} total {1e3 * (w4 - w0):.3f}")
    print(mode, "wall per step", 1e3 * (time.perf_counter() - t_all0) / steps, "ms")
    print(mode, {k: (v[0], round(v[1] / v[0], 3)) for k, v in _lib.prof_read().items()})
    _lib.prof_enable(False)
