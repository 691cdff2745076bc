#!/usr/bin/env python
"""Diagnostic: run one eval forward and dump the producer / consumer flags of the overlapped projection.

    python scripts/dump_sync.py B T [B T ...]
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

args = [int(a) for a in sys.argv[1:]] or [50, 2000]
for B, T in zip(args[::2], args[1::2]):
    torch.manual_seed(0)
    m = HeartSoundSegmenter(input_size=44, batch_size=B).eval()
    x = torch.randn(B, T, 44, device="cuda")
    torch.cuda.synchronize()
    t0 = time.time()
    logp, labels = m.forward_with_labels(x)
    torch.cuda.synchronize()
    dt = time.time() - t0
    ws = next(iter(m._workspace.values()))
    off = _lib.lib().hssb_debug_sync_offset(B, T)
    head = ws[off:off + 256].view(torch.int32).cpu()
    Q = 2 * ((T + 127) // 128)
    done = ws[off + 256:off + 256 + 4 * Q].view(torch.int32).cpu()
    tile = ws[off + 256 + 4 * 8192:off + 256 + 4 * 8192 + 4 * Q].view(torch.int32).cpu()
    print(f"B {B} T {T}: {dt * 1e3:.1f} ms, nan {bool(torch.isnan(logp).any())}, next_item {head[:8].tolist()}, timeout {int(head[8])}, resident {int(head[9])}")
    print("   chunk_done", done.tolist(), "need", B * 6 * 4)
    print("   tile_done ", tile.tolist())
    _lib.prof_enable(True); _lib.prof_read()
    m.forward_with_labels(x); torch.cuda.synchronize()
    print("   ", {k: round(v[1], 3) for k, v in _lib.prof_read().items()})
    _lib.prof_enable(False)
