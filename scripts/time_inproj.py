#!/usr/bin/env python
"""Per-kernel times of one BiLSTM forward (B=512, T=2000) -- used with the HSSB_IP_DEBUG / HSSB_RC_DEBUG knock-out switches
(those need a profiling build: HSSB_KNOCKOUTS=1 python __graft_entry__.py --force)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heart-sounds-segmentation_b200"))
sys.path.insert(0, ROOT)
import torch
from hss import _lib
from hss.model.segmenter import HeartSoundSegmenter

B, T = 512, int(sys.argv[1]) if len(sys.argv) > 1 else 2000
torch.manual_seed(0)
m = HeartSoundSegmenter(input_size=44, batch_size=B).eval()
x = torch.randn(B, T, 44, device="cuda")
m.predict(x)
_lib.prof_enable(True); _lib.prof_read()
for _ in range(3):
    m.predict(x)
torch.cuda.synchronize()
print({k: round(v[1] / v[0], 3) for k, v in _lib.prof_read().items()})
