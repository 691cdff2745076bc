"""Framing of a recording into fixed-length windows.

``frame_signal`` keeps the contract of reference ``hss/utils/preprocess.py:7-58`` -- ``L =
floor((T - n) / stride)`` frames of ``n`` rows starting every ``stride`` samples, one truncated
frame ``x[:n]`` when ``L <= 0`` -- and returns the same two lists of ``[n, C]`` views.

``frame_batch`` is the batched form the B200 path wants: one strided view ``[L, n]`` of the whole
recording (a single ``unfold``, no Python loop, works on CUDA tensors), ready for ``FSST.batch`` --
the 33 per-frame FFI calls of reference ``hss/datasets/heart_sounds.py:157-169`` become one.
"""
from __future__ import annotations

from typing import List, Tuple

import torch


def _num_frames(T: int, stride: int, n: int) -> int:
    if stride <= 0 or n <= 0:
        raise ValueError(f"stride and n must be positive, got stride={stride} n={n}")
    return (T - n) // stride if T >= n else -1     # floor((T - n) / stride) as the reference computes it


def frame_signal(x: torch.Tensor, y: torch.Tensor, stride: int, n: int) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
    """Frames of ``x`` and of its labels ``y`` (reference preprocess.py:7-58): lists of ``[n, C]`` views."""
    if x.dim() == 1:
        x = x.unsqueeze(1)
    if y.dim() == 1:
        y = y.unsqueeze(1)
    if x.shape[0] != y.shape[0]:
        raise AssertionError("signal and labels differ in length")
    L = _num_frames(int(x.shape[0]), stride, n)
    if L <= 0:
        return [x[:n, :]], [y[:n, :]]
    starts = range(0, L * stride, stride)
    return [x[s:s + n, :] for s in starts], [y[s:s + n, :] for s in starts]


def frame_batch(x: torch.Tensor, stride: int, n: int) -> torch.Tensor:
    """The same frames of a 1-D signal (or ``[T, 1]``) as one ``[L, n]`` tensor (strided view, no copy).

    Only defined when every frame is complete (``T >= n``); the truncated single frame of the
    reference's ``L <= 0`` branch is returned as ``[1, min(T, n)]``.
    """
    if x.dim() == 2 and x.shape[1] == 1:
        x = x[:, 0]
    if x.dim() != 1:
        raise ValueError(f"frame_batch expects a 1-D signal or [T, 1], got {tuple(x.shape)}")
    L = _num_frames(int(x.shape[0]), stride, n)
    if L <= 0:
        return x[:n].unsqueeze(0)
    return x.unfold(0, n, stride)[:L]
