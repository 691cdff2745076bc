import sys, time
import torch
sys.path.insert(0, "heart-sounds-segmentation_b200")
from hss.model.segmenter import HeartSoundSegmenter
from hss.optim import ClipAdam
B, T, F = 50, 2000, 44
torch.manual_seed(68)
m = HeartSoundSegmenter(input_size=F, batch_size=B).cuda().train()
opt = ClipAdam(m.parameters(), lr=0.01, max_norm=1.0)
x = torch.randn(B, T, F, device="cuda"); y = torch.randint(0, 4, (B, T), device="cuda")
def sync(): torch.cuda.synchronize()
def step(tm=None):
    def mark(k):
        if tm is not None:
            sync(); now = time.perf_counter(); tm[k] = tm.get(k, 0) + now - mark.t; mark.t = now
    mark.t = time.perf_counter()
    opt.zero_grad(set_to_none=True)
    h = m._packed(x.device); mark("repack")
    loss, _ = m.training_loss(x, y); mark("forward")
    loss.backward(); mark("backward")
    opt.step(); mark("optimizer")
for _ in range(3): step()
sync()
tm = {}
for _ in range(5): step(tm)
print({k: round(v / 5 * 1e3, 2) for k, v in tm.items()}, "sum", round(sum(tm.values()) / 5 * 1e3, 2))
sync(); t0 = time.perf_counter()
for _ in range(5): step()
sync(); print("unsynced step", round((time.perf_counter() - t0) / 5 * 1e3, 2))
# CPU-side enqueue time of each phase (no sync): how far the host runs ahead
t = {}
for _ in range(5):
    a = time.perf_counter(); opt.zero_grad(); m._packed(x.device); b = time.perf_counter()
    loss, _ = m.training_loss(x, y); c = time.perf_counter()
    loss.backward(); d = time.perf_counter(); opt.step(); e = time.perf_counter()
    for k, v in (("repack", b - a), ("forward", c - b), ("backward", d - c), ("optimizer", e - d)): t[k] = t.get(k, 0) + v
sync()
print("host enqueue", {k: round(v / 5 * 1e3, 2) for k, v in t.items()})
