#!/usr/bin/env python
"""Kernel-level breakdown of one training step at 50 x 2000 (torch.profiler: library kernels and torch's own side by side)."""
import sys
import torch
sys.path.insert(0, "heart-sounds-segmentation_b200")
from hss.model.segmenter import HeartSoundSegmenter
from hss.optim import ClipAdam

B, T, F = 50, 2000, 44
torch.manual_seed(68)
m = HeartSoundSegmenter(input_size=F, batch_size=B).cuda().train()
opt = ClipAdam(m.parameters(), lr=0.01, max_norm=1.0)
x = torch.randn(B, T, F, device="cuda")
y = torch.randint(0, 4, (B, T), device="cuda")


def step():
    opt.zero_grad(set_to_none=True)
    loss, _ = m.training_loss(x, y)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total / 3e3) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type == torch.autograd.DeviceType.CUDA]
rows.sort(key=lambda r: -r[2])
total = sum(r[2] for r in rows)
print(f"device time per step {total:.2f} ms over {sum(r[1] for r in rows) // 3} kernels")
for k, c, ms in rows[:28]:
    print(f"{ms:8.3f} ms  x{c // 3:<3d} {k[:130]}")
