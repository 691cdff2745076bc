"""Data-parallel plumbing: windows shard across ranks, the only collective is the metric all-reduce.

Every window / recording is independent in FSST (per-window z-score, reference
hss/transforms/synchrosqueeze.py:78-81) and in the BiLSTM (batch rows, reference
hss/model/segmenter.py:38-41), so ranks own contiguous blocks of windows and exchange nothing
until the 4x4 confusion counts (reference main.py:36-62, torchmetrics state) are summed.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block ``[lo, hi)`` of ``n`` windows owned by ``rank`` (sizes differ by at most 1)."""
    if world < 1 or not (0 <= rank < world) or n < 0:
        raise ValueError(f"bad shard request n={n} rank={rank} world={world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def confusion_counts(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """4x4 int64 counts ``cm[target, pred]`` of this rank's labels (CUDA kernel ``hssb_confusion``)."""
    if not pred.is_cuda:
        raise RuntimeError("confusion_counts runs on the GPU (no CPU fallback)")
    pred = pred.to(torch.int32).contiguous().reshape(-1)
    target = target.to(device=pred.device, dtype=torch.int64).contiguous().reshape(-1)
    if pred.numel() != target.numel():
        raise ValueError("pred / target size mismatch")
    cm = torch.zeros(16, dtype=torch.int64, device=pred.device)
    with torch.cuda.device(pred.device):
        rc = _lib.lib().hssb_confusion(pred.data_ptr(), target.data_ptr(), pred.numel(), cm.data_ptr(), _lib.stream_ptr())
    _lib.check(rc, "hssb_confusion")
    return cm.reshape(4, 4)


def allreduce_counts(cm: torch.Tensor) -> torch.Tensor:
    """Sum the per-rank counters over the job (NCCL on GPUs, gloo in CPU tests); no-op single-process."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(cm, op=dist.ReduceOp.SUM)
    return cm


def metric_state(logp: torch.Tensor, target: torch.Tensor, labels: torch.Tensor | None = None,
                 state: torch.Tensor | None = None) -> torch.Tensor:
    """The whole metric state of an evaluation step as 18 float64 scalars (CUDA kernel ``hssb_metrics_update``):
    ``state[:16]`` = confusion counts ``cm[target, pred]``, ``state[16]`` = summed loss terms, ``state[17]`` = elements.

    The loss is what the reference logs as ``train_loss`` / ``val_loss`` / ``test_loss`` (main.py:69-70,91-92,112-117):
    ``nn.CrossEntropyLoss`` on the permuted log-probabilities.  Pass ``state`` to accumulate over several steps on the
    device; ``allreduce_counts`` merges ranks with one all-reduce; ``metrics_from_state`` finalises.
    """
    if not logp.is_cuda:
        raise RuntimeError("metric_state runs on the GPU (no CPU fallback)")
    if logp.shape[-1] != 4:
        raise ValueError(f"expected [..., 4] log-probabilities, got {tuple(logp.shape)}")
    logp = logp.to(torch.float32).contiguous().reshape(-1, 4)
    target = target.to(device=logp.device, dtype=torch.int64).contiguous().reshape(-1)
    if logp.shape[0] != target.numel():
        raise ValueError("logp / target size mismatch")
    if labels is not None:
        labels = labels.to(device=logp.device, dtype=torch.int32).contiguous().reshape(-1)
        if labels.numel() != target.numel():
            raise ValueError("labels / target size mismatch")
    if state is None:
        state = torch.zeros(18, dtype=torch.float64, device=logp.device)
    elif state.shape != (18,) or state.dtype != torch.float64 or state.device != logp.device or not state.is_contiguous():
        raise ValueError("state must be a contiguous float64 [18] tensor on the device of logp")
    with torch.cuda.device(logp.device):
        rc = _lib.lib().hssb_metrics_update(logp.data_ptr(), labels.data_ptr() if labels is not None else None, target.data_ptr(),
                                            target.numel(), state.data_ptr(), _lib.stream_ptr())
    _lib.check(rc, "hssb_metrics_update")
    return state


def metrics_from_state(state: torch.Tensor) -> dict:
    """Metrics of ``metrics_from_counts`` plus ``loss`` (mean cross-entropy) from an (all-reduced) 18-scalar state."""
    st = state.detach().to(torch.float64).cpu()
    out = metrics_from_counts(st[:16].round().to(torch.int64).reshape(4, 4))
    out["loss"] = float(st[16] / st[17]) if float(st[17]) > 0 else float("nan")
    out["count"] = int(st[17].round())
    return out


def metrics_from_counts(cm: torch.Tensor) -> dict:
    """Per-class and macro accuracy(=recall) / precision / F1 from ``cm[target, pred]``.

    Same definitions as the torchmetrics multiclass collection of reference main.py:36-62
    (``Accuracy(average=None)`` is per-class recall; zero-division -> 0).  ``average="macro"`` follows torchmetrics'
    ``_adjust_weights_safe_divide``: a class that occurs neither in the targets nor in the predictions
    (tp + fp + fn == 0) gets weight 0, so the macro figures are means over the classes that are present.
    """
    cm = cm.to(torch.float64).cpu()
    tp = cm.diag()
    support = cm.sum(dim=1)
    predicted = cm.sum(dim=0)
    recall = torch.where(support > 0, tp / support.clamp(min=1), torch.zeros_like(tp))
    precision = torch.where(predicted > 0, tp / predicted.clamp(min=1), torch.zeros_like(tp))
    denom = precision + recall
    f1 = torch.where(denom > 0, 2 * precision * recall / denom.clamp(min=1e-300), torch.zeros_like(tp))
    present = ((support + predicted) > 0).to(torch.float64)
    n_present = present.sum().clamp(min=1)

    def macro(v):
        return float((v * present).sum() / n_present)

    return {
        "accuracy_per_class": recall, "recall_per_class": recall, "precision_per_class": precision, "f1_per_class": f1,
        "accuracy": macro(recall), "recall": macro(recall), "precision": macro(precision),
        "f1": macro(f1), "micro_accuracy": float(tp.sum() / cm.sum().clamp(min=1)),
    }


def score_histograms(logp: torch.Tensor, target: torch.Tensor, nbins: int = 4096) -> torch.Tensor:
    """One-vs-rest score histograms ``hist[class, is_positive, bin]`` (int64) of this rank's log-probabilities.

    State of the binned multiclass AUROC (reference main.py:48,60); plain counters, so
    ``allreduce_counts`` merges ranks exactly like the confusion counts.  CUDA kernel ``hssb_auroc_hist``.
    """
    if not logp.is_cuda:
        raise RuntimeError("score_histograms runs on the GPU (no CPU fallback)")
    if logp.shape[-1] != 4:
        raise ValueError(f"expected [..., 4] log-probabilities, got {tuple(logp.shape)}")
    logp = logp.to(torch.float32).contiguous().reshape(-1, 4)
    target = target.to(device=logp.device, dtype=torch.int64).contiguous().reshape(-1)
    if logp.shape[0] != target.numel():
        raise ValueError("logp / target size mismatch")
    hist = torch.zeros(4, 2, nbins, dtype=torch.int64, device=logp.device)
    with torch.cuda.device(logp.device):
        rc = _lib.lib().hssb_auroc_hist(logp.data_ptr(), target.data_ptr(), target.numel(), nbins, hist.data_ptr(), _lib.stream_ptr())
    _lib.check(rc, "hssb_auroc_hist")
    return hist


def auroc_from_histograms(hist: torch.Tensor) -> dict:
    """Per-class and macro one-vs-rest AUROC from (all-reduced) score histograms.

    Area under the ROC curve whose thresholds are the bin edges, trapezoidal rule (scores sharing a bin
    count as ties) -- the Mann-Whitney form ``sum_b pos[b] * (neg_below[b] + neg[b] / 2) / (P * N)``.
    A class without positives or without negatives scores 0 (torchmetrics' convention for an undefined
    curve); macro = mean over the four classes (reference main.py:60).
    """
    h = hist.to(torch.float64).cpu()
    neg, pos = h[:, 0], h[:, 1]
    below = torch.cumsum(neg, dim=1) - neg
    num = (pos * (below + 0.5 * neg)).sum(dim=1)
    den = pos.sum(dim=1) * neg.sum(dim=1)
    per_class = torch.where(den > 0, num / den.clamp(min=1), torch.zeros_like(num))
    return {"auroc_per_class": per_class, "auroc": float(per_class.mean())}
