#!/usr/bin/env python
"""Generates ``tests/data/0001.csv``: the stand-in for one DavidSpringerHSS recording (BASELINE config 1).

The real dataset is a network download (reference hss/datasets/heart_sounds.py:136-151).  Same on-disk format as the
files ``_load_file`` reads (heart_sounds.py:193-197): a header row, column 0 the PCG signal (float), column 1 the state
labels 1..4.  35 000 samples at 1 kHz -> 33 frames of 2000 at stride 1000 (reference test/test_dataset.py:37).
Signal: seed 68, 0.05*N(0,1) + S1 / S2 bursts every 0.83 s (workloads.synth_pcg); labels cyclic 1,2,3,4 with durations
120/200/100/410 ms.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from workloads import synth_pcg, synthetic_targets  # noqa: E402

N = 35_000
x = synth_pcg(N, 1000.0, 68)
y = synthetic_targets(1, N)[0] + 1
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "0001.csv"), "w") as f:
    f.write("Signals,Labels\n")
    for a, b in zip(x, y):
        f.write(f"{float(a):.9g},{int(b)}\n")
