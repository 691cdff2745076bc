#!/usr/bin/env python
"""Workloads for compute-sanitizer (racecheck / synccheck / memcheck): every default kernel at sizes that span several
clusters, the overlapped projection (T >= 1024) and one training step through the cluster kernels.

    compute-sanitizer --tool racecheck python scripts/sanitize_r2.py [fsst] [lstm] [overlap] [train]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heart-sounds-segmentation_b200"))
sys.path.insert(0, ROOT)
import torch
from hss.model.segmenter import HeartSoundSegmenter
from hss.transforms import FSST
from workloads import reference_window, synth_pcg_batch

what = sys.argv[1:] or ["fsst", "lstm", "overlap", "pipeline", "train"]
if "fsst" in what:
    for B, N in ((3, 300), (37, 170)):
        x = torch.from_numpy(synth_pcg_batch(B, N)).cuda()
        feats = FSST(1000.0, window=reference_window(128), truncate_freq=(25, 200), stack=True).batch(x)
        torch.cuda.synchronize()
        print("fsst", B, N, float(feats.abs().mean()))
if "lstm" in what:
    for B, T in ((70, 24), (200, 12)):          # 3 and 7 sub-tiles per direction: several clusters, ragged last group
        torch.manual_seed(1)
        m = HeartSoundSegmenter(input_size=44, batch_size=B).eval()
        logp, labels = m.forward_with_labels(torch.randn(B, T, 44, device="cuda"))
        torch.cuda.synchronize()
        print("lstm", B, T, float(logp.exp().sum(-1).mean()), int(labels.sum()))
if "overlap" in what:
    B, T = 40, 1030                              # 9 time tiles: the projection runs as launches M / A / B around the recurrences
    os.environ.setdefault("HSSB_K4_MID", "25")
    torch.manual_seed(2)
    m = HeartSoundSegmenter(input_size=44, batch_size=B).eval()
    logp, labels = m.forward_with_labels(torch.randn(B, T, 44, device="cuda"))
    torch.cuda.synchronize()
    print("overlap", B, T, float(logp.exp().sum(-1).mean()), int(labels.sum()))
if "pipeline" in what:
    # hss.pipeline: the next batch's FSST behind hssb_model_side_gate on a second stream, the projection launch M behind its tile gate
    import numpy as np
    from hss.pipeline import SegmentationPipeline

    B, T = 40, 1030
    os.environ.setdefault("HSSB_K4_MID", "25")
    torch.manual_seed(4)
    fsst = FSST(1000.0, window=np.kaiser(128, 0.5), truncate_freq=(25, 200), stack=True)
    m = HeartSoundSegmenter(input_size=44, batch_size=B).eval()
    xs = [torch.from_numpy(synth_pcg_batch(B, T)).cuda() * s_ for s_ in (1.0, 0.5, 2.0)]
    pipe = SegmentationPipeline(fsst, m)
    tot = 0
    for i, x in enumerate(xs):
        logp, labels = pipe(x, prefetch=xs[i + 1] if i + 1 < len(xs) else None)
        tot += int(labels.sum())
    pipe.close()
    torch.cuda.synchronize()
    print("pipeline", B, T, tot)
if "train" in what or "train_simt" in what:
    # "train": the default tensor-core path (K4 + K5m TRAIN forward, K5b backward with 8 / 16 columns per cluster, TF32-split
    # GEMM operands, fused head + loss, clip + Adam); "train_simt": the fp32 cluster kernels (HSSB_TRAIN_IMPL=cluster)
    from hss.optim import ClipAdam

    if "train_simt" in what:
        os.environ["HSSB_TRAIN_IMPL"] = "cluster"
    # SAN_TRAIN_SHAPES="9x20,70x12": racecheck needs minutes per shape on the spinning cluster kernels
    shapes = [tuple(int(v) for v in s_.split("x")) for s_ in os.environ.get("SAN_TRAIN_SHAPES", "9x20,70x12").split(",")]
    for B, T in shapes:
        torch.manual_seed(3)
        m = HeartSoundSegmenter(input_size=44, batch_size=B).cuda().train()
        opt = ClipAdam(m.parameters(), lr=0.01, max_norm=1.0)
        x = torch.randn(B, T, 44, device="cuda")
        y = torch.randint(0, 4, (B, T), device="cuda")
        for _ in range(2):
            opt.zero_grad()
            loss, _ = m.training_loss(x, y)
            loss.backward()
            opt.step()
        torch.cuda.synchronize()
        print("train", os.environ.get("HSSB_TRAIN_IMPL", "tc"), B, T, float(loss))
