"""CPU oracle for the evaluation metrics -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.

Restates what reference ``main.py:36-62`` asks of torchmetrics (absent from this image; definitions
from its documentation, multiclass task, 4 classes):
  * ``confusion``        counts ``cm[target, pred]``
  * ``per_class``        Accuracy(average=None) = per-class recall, Precision, Recall, F1 with
                         zero-division -> 0; the ``macro`` variants are their plain means
  * ``auroc_exact``      AUROC(average=None / "macro"), thresholds=None: one-vs-rest area under the
                         exact ROC curve of the class probability (softmax of the model output; the
                         model emits log-softmax, so the probability is exp(logp)), ties sharing a
                         trapezoid = the Mann-Whitney U statistic with mid-ranks.  An undefined
                         curve (no positives or no negatives) scores 0.
  * ``auroc_binned``     the same on scores quantised to ``floor(p * nbins)`` -- what the product's
                         all-reducible histogram state represents.

Pinned: ``tests/test_host_logic.py`` checks ``auroc_exact`` against scikit-learn's ``roc_auc_score``
(an independent implementation of the same definition); parity with torchmetrics itself is unpinned.
"""
from __future__ import annotations

import numpy as np


def confusion(pred: np.ndarray, target: np.ndarray, classes: int = 4) -> np.ndarray:
    cm = np.zeros((classes, classes), np.int64)
    for t, p in zip(np.asarray(target).reshape(-1), np.asarray(pred).reshape(-1)):
        if 0 <= t < classes and 0 <= p < classes:
            cm[t, p] += 1
    return cm


def per_class(cm: np.ndarray) -> dict:
    cm = cm.astype(np.float64)
    tp = np.diag(cm)
    support, predicted = cm.sum(1), cm.sum(0)
    recall = np.divide(tp, support, out=np.zeros_like(tp), where=support > 0)
    precision = np.divide(tp, predicted, out=np.zeros_like(tp), where=predicted > 0)
    f1 = np.divide(2 * precision * recall, precision + recall, out=np.zeros_like(tp), where=(precision + recall) > 0)
    return {"accuracy": recall, "recall": recall, "precision": precision, "f1": f1}


def _mann_whitney(score: np.ndarray, positive: np.ndarray) -> float:
    n_pos = int(positive.sum())
    n_neg = positive.size - n_pos
    if n_pos == 0 or n_neg == 0:
        return 0.0
    order = np.argsort(score, kind="mergesort")
    s = score[order]
    ranks = np.empty(s.size, np.float64)
    i = 0
    while i < s.size:                      # mid-ranks over runs of equal scores
        j = i
        while j + 1 < s.size and s[j + 1] == s[i]:
            j += 1
        ranks[i:j + 1] = 0.5 * (i + j) + 1.0
        i = j + 1
    r_pos = ranks[positive[order]].sum()
    return float((r_pos - n_pos * (n_pos + 1) / 2.0) / (n_pos * n_neg))


def auroc_exact(logp: np.ndarray, target: np.ndarray) -> np.ndarray:
    p = np.exp(np.asarray(logp, np.float32).reshape(-1, 4))
    t = np.asarray(target).reshape(-1)
    keep = (t >= 0) & (t < 4)
    return np.array([_mann_whitney(p[keep, c], t[keep] == c) for c in range(4)])


def auroc_binned(logp: np.ndarray, target: np.ndarray, nbins: int) -> np.ndarray:
    p = np.exp(np.asarray(logp, np.float32).reshape(-1, 4))
    q = np.minimum(nbins - 1, np.floor(p * np.float32(nbins)).astype(np.int64))
    t = np.asarray(target).reshape(-1)
    keep = (t >= 0) & (t < 4)
    return np.array([_mann_whitney(q[keep, c].astype(np.float64), t[keep] == c) for c in range(4)])


def histograms(logp: np.ndarray, target: np.ndarray, nbins: int) -> np.ndarray:
    p = np.exp(np.asarray(logp, np.float32).reshape(-1, 4))
    q = np.minimum(nbins - 1, np.floor(p * np.float32(nbins)).astype(np.int64))
    t = np.asarray(target).reshape(-1)
    h = np.zeros((4, 2, nbins), np.int64)
    for c in range(4):
        for pos in (0, 1):
            sel = (t >= 0) & (t < 4) & ((t == c) == bool(pos))
            h[c, pos] = np.bincount(q[sel, c], minlength=nbins)
    return h
