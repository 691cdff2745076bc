"""The N>1 host path on CPU: world_size-2 gloo job shards the windows and all-reduces the metric
counters; the summed confusion matrix must equal the single-process one."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cpu_counts(pred, target):
    cm = torch.zeros(4, 4, dtype=torch.int64)
    for t, p in zip(target.reshape(-1).tolist(), pred.reshape(-1).tolist()):
        cm[t, p] += 1
    return cm


def _worker(rank, world, port, pred, target, out_path):
    sys.path.insert(0, os.path.join(ROOT, "heart-sounds-segmentation_b200"))
    from hss.sharding import allreduce_counts, shard_range

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = shard_range(pred.shape[0], rank, world)
    cm = _cpu_counts(pred[lo:hi], target[lo:hi])
    cm = allreduce_counts(cm)
    if rank == 0:
        torch.save({"cm": cm, "range": (lo, hi)}, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_confusion_allreduce(tmp_path):
    g = torch.Generator().manual_seed(0)
    pred = torch.randint(0, 4, (7, 50), generator=g)
    target = torch.randint(0, 4, (7, 50), generator=g)
    out = str(tmp_path / "cm.pt")
    mp.spawn(_worker, args=(2, _free_port(), pred, target, out), nprocs=2, join=True)
    got = torch.load(out)
    assert torch.equal(got["cm"], _cpu_counts(pred, target))
    assert got["range"] == (0, 4)


def _hist_worker(rank, world, port, logp, target, out_path):
    sys.path.insert(0, os.path.join(ROOT, "heart-sounds-segmentation_b200"))
    sys.path.insert(0, ROOT)
    from hss.sharding import allreduce_counts, auroc_from_histograms, shard_range
    from oracle import metrics_oracle as mo

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = shard_range(logp.shape[0], rank, world)
    # (on a GPU rank this histogram comes from hss.sharding.score_histograms = kernel hssb_auroc_hist)
    hist = torch.from_numpy(mo.histograms(logp[lo:hi].numpy(), target[lo:hi].numpy(), 256))
    hist = allreduce_counts(hist)
    if rank == 0:
        torch.save({"hist": hist, "auroc": auroc_from_histograms(hist)}, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_auroc_histogram_allreduce(tmp_path):
    """The AUROC state (score histograms, reference main.py:48,60) shards like the confusion counts: the all-reduced
    histogram of two ranks equals the single-process one, and so does the AUROC computed from it."""
    sys.path.insert(0, os.path.join(ROOT, "heart-sounds-segmentation_b200"))
    from hss.sharding import auroc_from_histograms

    from oracle import metrics_oracle as mo

    g = torch.Generator().manual_seed(1)
    target = torch.randint(0, 4, (9, 40), generator=g)
    logits = torch.randn(9, 40, 4, generator=g) + 1.5 * torch.nn.functional.one_hot(target, 4)
    logp = torch.log_softmax(logits, dim=2)
    out = str(tmp_path / "hist.pt")
    mp.spawn(_hist_worker, args=(2, _free_port(), logp, target, out), nprocs=2, join=True)
    got = torch.load(out)
    whole = torch.from_numpy(mo.histograms(logp.numpy(), target.numpy(), 256))
    assert torch.equal(got["hist"], whole)
    assert torch.equal(got["auroc"]["auroc_per_class"], auroc_from_histograms(whole)["auroc_per_class"])
    assert abs(got["auroc"]["auroc"] - float(mo.auroc_binned(logp.numpy(), target.numpy(), 256).mean())) < 1e-12


def _state_worker(rank, world, port, logp, target, out_path):
    sys.path.insert(0, os.path.join(ROOT, "heart-sounds-segmentation_b200"))
    from hss.sharding import allreduce_counts, metrics_from_state, shard_range

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = shard_range(logp.shape[0], rank, world)
    # (on a GPU rank this state comes from hss.sharding.metric_state = kernel hssb_metrics_update, accumulated over the steps)
    lp, t = logp[lo:hi].reshape(-1, 4).double(), target[lo:hi].reshape(-1)
    state = torch.zeros(18, dtype=torch.float64)
    state[:16] = _cpu_counts(lp.argmax(-1), t).reshape(-1).double()
    state[16] = (torch.logsumexp(lp, dim=1) - lp[torch.arange(lp.shape[0]), t]).sum()
    state[17] = lp.shape[0]
    state = allreduce_counts(state)                      # the ONE collective of the job
    if rank == 0:
        torch.save({"state": state, "metrics": metrics_from_state(state)}, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_metric_state_allreduce(tmp_path):
    """The 18-scalar metric state (confusion counts + loss sum + count, reference main.py:36-62,69-70) of two ranks adds up to
    the single-process state: same confusion matrix, same mean cross-entropy as nn.CrossEntropyLoss on the whole batch."""
    g = torch.Generator().manual_seed(4)
    logp = torch.log_softmax(torch.randn(9, 40, 4, generator=g), dim=2)
    target = torch.randint(0, 4, (9, 40), generator=g)
    out = str(tmp_path / "state.pt")
    mp.spawn(_state_worker, args=(2, _free_port(), logp, target, out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    assert torch.equal(got["state"][:16].round().long().reshape(4, 4), _cpu_counts(logp.argmax(-1), target))
    ref_loss = torch.nn.CrossEntropyLoss()(logp.permute(0, 2, 1).double(), target)
    assert abs(got["metrics"]["loss"] - float(ref_loss)) < 1e-12 and got["metrics"]["count"] == 9 * 40
