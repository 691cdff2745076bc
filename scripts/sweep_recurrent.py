#!/usr/bin/env python
"""Time the BiLSTM forward for several batch sizes and recurrence geometries (HSSB_RC_GEOM values).

    python scripts/sweep_recurrent.py "50,128,256,512" "auto;32,1,2;32,2,2;32,3,2" [T]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heart-sounds-segmentation_b200"))
sys.path.insert(0, ROOT)
import torch
from hss import _lib
from hss.model.segmenter import HeartSoundSegmenter

Bs = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "50,128,256,512").split(",")]
geoms = (sys.argv[2] if len(sys.argv) > 2 else "auto;32,1,2;32,2,2;32,3,2").split(";")
T = int(sys.argv[3]) if len(sys.argv) > 3 else 500
for B in Bs:
    torch.manual_seed(0)
    m = HeartSoundSegmenter(input_size=44, batch_size=B).eval()
    x = torch.randn(B, T, 44, device="cuda")
    ref = None
    for g in geoms:
        if g == "auto":
            os.environ.pop("HSSB_RC_GEOM", None)
        else:
            os.environ["HSSB_RC_GEOM"] = g
        try:
            lab = m.predict(x)
            _lib.prof_enable(True); _lib.prof_read()
            for _ in range(3):
                lab = m.predict(x)
            torch.cuda.synchronize()
            prof = _lib.prof_read()
            _lib.prof_enable(False)
        except Exception as e:  # noqa: BLE001
            print(f"B {B:4d} geom {g:8s}: failed: {e}")
            continue
        if ref is None:
            ref = lab
        same = bool(torch.equal(ref, lab))
        rc = prof["tc_recurrent"]
        print(f"B {B:4d} T {T} geom {g:8s}: tc_recurrent {rc[1] / 3:8.3f} ms/forward ({rc[0] // 3} launches)  labels equal to first geometry: {same}", flush=True)
