#!/usr/bin/env python
"""Row f2 timing: the native batch CSV parser + pipelined ingest against the reference's pandas loop.

    python scripts/time_ingest.py [n_files] [rows]

Writes n_files synthetic recordings (two-column CSV, the DavidSpringerHSS on-disk format) to a temporary directory and times
  (a) pandas:   pd.read_csv per file -> torch tensors            (reference hss/datasets/heart_sounds.py:193-197)
  (b) native:   hssb_csv_scan + hssb_csv_parse, all files, host thread pool -> pinned staging
  (c) reference-style loop on the GPU: pandas per file -> recording_to_frames (H2D + FSST.frames) per file
  (d) stream_recordings: native parse of group k+1 overlapped with H2D + FSST of group k
"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heart-sounds-segmentation_b200"))
sys.path.insert(0, ROOT)
import numpy as np
import pandas as pd
import torch
from hss.transforms import FSST
from hss.utils import load_recordings_csv, recording_to_frames, stream_recordings
from workloads import reference_window, synth_pcg, synthetic_targets

n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 35_000
tmp = tempfile.mkdtemp(prefix="hssb_ingest_")
paths = []
base = synth_pcg(rows, 1000.0, 68)
y = synthetic_targets(1, rows)[0] + 1
for i in range(n_files):
    p = os.path.join(tmp, f"{i:04d}.csv")
    x = base * (0.5 + i / n_files)
    with open(p, "w") as f:
        f.write("Signals,Labels\n")
        f.write("\n".join(f"{float(a):.9g},{int(b)}" for a, b in zip(x, y)))
        f.write("\n")
    paths.append(p)
mb = sum(os.path.getsize(p) for p in paths) / 1e6
samples = n_files * rows
print(f"{n_files} files x {rows} rows = {mb:.1f} MB of CSV, {os.cpu_count()} host cores")


def pandas_load(p):
    df = pd.read_csv(p, skiprows=1, names=["Signals", "Labels"])
    return torch.tensor(df.loc[:, "Signals"].to_numpy(), dtype=torch.float32), torch.tensor(df.loc[:, "Labels"].to_numpy(), dtype=torch.int64)


t0 = time.perf_counter(); ref = [pandas_load(p) for p in paths]; t_pandas = time.perf_counter() - t0
from hss.utils.ingest import _staging

load_recordings_csv(paths[:2])
t0 = time.perf_counter(); staging = _staging(samples, torch.cuda.is_available()); t_pin = time.perf_counter() - t0    # one-time pinned allocation
t0 = time.perf_counter(); rec = load_recordings_csv(paths, pin=torch.cuda.is_available(), staging=staging); t_native = time.perf_counter() - t0
assert all(torch.equal(rec[i][0], ref[i][0]) and torch.equal(rec[i][1], ref[i][1]) for i in range(n_files))
print(f"(a) pandas loop        {t_pandas * 1e3:8.1f} ms  {samples / t_pandas / 1e6:7.2f} M samples/s  {mb / t_pandas:7.1f} MB/s")
print(f"    (one-time allocation of the pinned staging buffers: {t_pin * 1e3:.1f} ms)")
print(f"(b) native batch parse {t_native * 1e3:8.1f} ms  {samples / t_native / 1e6:7.2f} M samples/s  {mb / t_native:7.1f} MB/s   ({t_pandas / t_native:.1f}x, values bit-equal)")
if torch.cuda.is_available():
    fsst = FSST(1000.0, window=reference_window(128), truncate_freq=(25, 200), stack=True)
    recording_to_frames(*pandas_load(paths[0]), fsst)
    list(stream_recordings(paths[:2], fsst))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n_frames = 0
    for p in paths:
        feats, labels = recording_to_frames(*pandas_load(p), fsst)
        n_frames += feats.shape[0]
    torch.cuda.synchronize(); t_loop = time.perf_counter() - t0
    t0 = time.perf_counter()
    n2 = sum(f.shape[0] for f, _ in stream_recordings(paths, fsst, group=16))
    torch.cuda.synchronize(); t_stream = time.perf_counter() - t0
    assert n2 == n_frames
    print(f"(c) pandas + per-file FSST.frames   {t_loop * 1e3:8.1f} ms  {n_frames / t_loop:8.0f} frames/s")
    print(f"(d) stream_recordings (pipelined)   {t_stream * 1e3:8.1f} ms  {n2 / t_stream:8.0f} frames/s   ({t_loop / t_stream:.1f}x)")
