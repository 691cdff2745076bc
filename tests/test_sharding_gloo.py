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
