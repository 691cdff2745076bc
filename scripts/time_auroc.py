import sys, torch
sys.path.insert(0, "heart-sounds-segmentation_b200")
from hss.sharding import score_histograms
from hss import _lib
n = 512 * 2000
g = torch.Generator().manual_seed(0)
target = torch.randint(0, 4, (n,), generator=g).cuda()
for sharp in (0.0, 20.0):
    logp = torch.log_softmax(torch.randn(n, 4, generator=g).cuda() + sharp * torch.nn.functional.one_hot(target, 4), dim=1)
    for _ in range(3): score_histograms(logp, target)
    _lib.prof_enable(True)
    for _ in range(10): score_histograms(logp, target)
    torch.cuda.synchronize()
    print("sharp", sharp, _lib.prof_read())
    _lib.prof_enable(False)
