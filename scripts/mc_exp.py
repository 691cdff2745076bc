import os, sys
sys.path.insert(0, "heart-sounds-segmentation_b200"); sys.path.insert(0, ".")
import torch
from hss.model.segmenter import HeartSoundSegmenter
B, T = int(sys.argv[1]), int(sys.argv[2])
torch.manual_seed(1)
m = HeartSoundSegmenter(input_size=44, batch_size=B).eval()
logp, labels = m.forward_with_labels(torch.randn(B, T, 44, device="cuda"))
torch.cuda.synchronize()
print("ok", B, T, float(logp.exp().sum(-1).mean()))
