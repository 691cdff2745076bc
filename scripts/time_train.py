#!/usr/bin/env python
"""Time of one training step (forward in training mode + backward + clip + Adam) at BASELINE config 3's size
(batch 50, T 2000, 44 features), per-kernel split from the library's launch timers."""
import sys, time
import torch
sys.path.insert(0, "heart-sounds-segmentation_b200")
from hss import _lib
from hss.model.segmenter import HeartSoundSegmenter

B, T, F = 50, 2000, 44
torch.manual_seed(68)
m = HeartSoundSegmenter(input_size=F, batch_size=B).cuda().train()
opt = torch.optim.Adam(m.parameters(), lr=0.01)
x = torch.randn(B, T, F, device="cuda")
y = torch.randint(0, 4, (B, T), device="cuda")


from hss.optim import ClipAdam

fused = "--eager" not in sys.argv          # fused head + loss and clip + Adam kernels (default) or the reference's eager torch ops
if fused:
    opt = ClipAdam(m.parameters(), lr=0.01, max_norm=1.0)


def step():
    opt.zero_grad(set_to_none=True)
    if fused:
        loss, _ = m.training_loss(x, y)
        loss.backward()
    else:
        loss = torch.nn.functional.cross_entropy(m(x).permute(0, 2, 1), y)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
    opt.step()
    return loss


for _ in range(8):          # (the first asynchronous steps grow torch's allocator cache: one-off cudaMallocs of several GB)
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
print(f"training step, launch timers off: {(time.perf_counter() - t0) / 5 * 1e3:.1f} ms")
_lib.prof_enable(True)
_lib.prof_read()
t0 = time.perf_counter()
n = 3
for _ in range(n):
    loss = step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
prof = _lib.prof_read()
print(f"with the per-launch timers on ({'fused head/loss + clip/Adam kernels' if fused else 'eager torch head / loss / clip / Adam'}) B={B} T={T}: {dt * 1e3:.1f} ms  ({B * T / dt / 1e6:.2f} M samples/s), loss {float(loss.detach()):.4f}")
for k, (c, ms) in sorted(prof.items()):
    print(f"  {k}: {c // n} launches/step, {ms / n:.2f} ms/step")
