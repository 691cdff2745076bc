import sys, time
sys.path.insert(0, "heart-sounds-segmentation_b200")
import torch
from hss.model.segmenter import HeartSoundSegmenter
from hss.optim import ClipAdam
B, T, F = 50, 2000, 44
torch.manual_seed(68)
m = HeartSoundSegmenter(input_size=F, batch_size=B).cuda().train()
opt = ClipAdam(m.parameters(), lr=0.01, max_norm=1.0)
x = torch.randn(B, T, F, device="cuda"); y = torch.randint(0, 4, (B, T), device="cuda")
def step():
    opt.zero_grad(set_to_none=True)
    loss, _ = m.training_loss(x, y); loss.backward(); opt.step()
for mode in ("sync each", "async x5", "sync each", "async x5", "async x5"):
    ts = []
    if mode == "sync each":
        for _ in range(6):
            torch.cuda.synchronize(); t0 = time.perf_counter(); step(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    else:
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5):
            a = time.perf_counter(); step(); ts.append((time.perf_counter() - a) * 1e3)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3 / 5)
    print(mode, " ".join(f"{t:.1f}" for t in ts), "| reserved GB", round(torch.cuda.memory_reserved() / 2**30, 2), flush=True)
