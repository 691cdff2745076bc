import os, sys, time
sys.path.insert(0, "heart-sounds-segmentation_b200"); sys.path.insert(0, ".")
import torch
from hss.model.segmenter import HeartSoundSegmenter
from hss.optim import ClipAdam
B, T = int(sys.argv[1]), int(sys.argv[2])
torch.manual_seed(3)
m = HeartSoundSegmenter(input_size=44, batch_size=B).cuda().train()
opt = ClipAdam(m.parameters(), lr=0.01, max_norm=1.0)
x = torch.randn(B, T, 44, device="cuda"); y = torch.randint(0, 4, (B, T), device="cuda")
t0 = time.time()
def mark(s):
    torch.cuda.synchronize(); print(f"{s} {time.time() - t0:.1f}s", flush=True)
for it in range(2):
    opt.zero_grad(); mark(f"step {it} start")
    loss, _ = m.training_loss(x, y); mark("forward")
    loss.backward(); mark("backward")
    opt.step(); mark("optimizer")
print("ok", B, T, float(loss.detach()))
