#!/usr/bin/env python
"""Per-kernel times of FSST.batch (1024 windows x 2000 samples: BASELINE config 2)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heart-sounds-segmentation_b200"))
sys.path.insert(0, ROOT)
import torch
from hss import _lib
from hss.transforms import FSST
from workloads import reference_window, tiled_windows

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x = torch.from_numpy(tiled_windows(B, 2000, 1000.0, 68)).cuda()
f = FSST(1000.0, window=reference_window(128), truncate_freq=(25, 200), stack=True)
f.batch(x)
_lib.prof_enable(True); _lib.prof_read()
for _ in range(5):
    f.batch(x)
torch.cuda.synchronize()
prof = _lib.prof_read()
tot = sum(v[1] / v[0] for v in prof.values())
print({k: round(v[1] / v[0], 4) for k, v in prof.items()}, "total ms", round(tot, 4), "M samples/s", round(B * 2000 / tot / 1e3, 1))
