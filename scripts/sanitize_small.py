#!/usr/bin/env python
"""Tiny FSST + BiLSTM forwards for compute-sanitizer (memcheck): ragged batch / time sizes through every default kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heart-sounds-segmentation_b200"))
sys.path.insert(0, ROOT)
import torch
from hss.model.segmenter import HeartSoundSegmenter
from hss.transforms import FSST
from workloads import reference_window, synth_pcg_batch

for B, N in ((3, 300), (37, 170)):
    x = torch.from_numpy(synth_pcg_batch(B, N)).cuda()
    feats = FSST(1000.0, window=reference_window(128), truncate_freq=(25, 200), stack=True).batch(x)
    torch.manual_seed(1)
    m = HeartSoundSegmenter(input_size=44, batch_size=B).eval()
    logp, labels = m.forward_with_labels(feats)
    torch.cuda.synchronize()
    print(B, N, float(logp.exp().sum(-1).mean()), int(labels.sum()))
os.environ["HSSB_FUSE_X"] = "0"
B, N = 5, 140
feats = torch.randn(B, N, 44, device="cuda")
m = HeartSoundSegmenter(input_size=44, batch_size=B).eval()
print(float(m(feats).exp().sum(-1).mean()))
torch.cuda.synchronize()
