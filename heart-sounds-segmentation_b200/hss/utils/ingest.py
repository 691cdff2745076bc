"""Recording ingest for the B200 path: CSV -> pinned host tensors -> framed, transformed batch.

``load_recording_csv`` keeps the contract of reference ``hss/datasets/heart_sounds.py:193-197`` (``_load_file``): a
two-column CSV whose first row is a header, column 0 the PCG signal, column 1 the labels 1..4.  ``recording_to_frames``
is the in-memory + framing branch of the dataset constructor (``heart_sounds.py:155-169``) as ONE device call: frames of
``frame_len`` samples every ``stride`` samples, labels shifted to 0..3 (``y - 1``, ``:164``), every frame transformed by
``FSST`` (``:166``) -- recordings shorter than one frame are skipped (``:160-161``).
"""
from __future__ import annotations

import numpy as np
import torch

from .preprocess import frame_batch, frame_signal


def load_recording_csv(path: str, dtype: torch.dtype = torch.float32, pin: bool = False) -> tuple[torch.Tensor, torch.Tensor]:
    """``(signal [T] dtype, labels [T] int64)`` of one recording file (header row skipped), optionally in pinned memory."""
    table = np.loadtxt(path, delimiter=",", skiprows=1, ndmin=2, dtype=np.float64)
    if table.shape[1] < 2:
        raise ValueError(f"{path}: expected two columns (signal, label), got {table.shape[1]}")
    x = torch.from_numpy(table[:, 0].copy()).to(dtype)
    y = torch.from_numpy(table[:, 1].astype(np.int64))
    if pin and torch.cuda.is_available():
        x, y = x.pin_memory(), y.pin_memory()
    return x, y


def recording_to_frames(x: torch.Tensor, y: torch.Tensor, fsst, stride: int = 1000, frame_len: int = 2000,
                        device: torch.device | str | None = None) -> tuple[torch.Tensor, torch.Tensor]:
    """Features ``[L, frame_len, F]`` and labels ``[L, frame_len]`` (0-based) of all frames of one recording.

    One H2D copy of the recording and one ``FSST.frames`` call replace the per-frame transform loop of the reference.
    Returns empty tensors for a recording shorter than ``frame_len``.
    """
    if x.dim() == 2 and x.shape[1] == 1:
        x = x[:, 0]
    if x.shape[0] != y.shape[0]:
        raise AssertionError("signal and labels differ in length")
    dev = torch.device(device) if device is not None else (x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device()))
    if x.shape[0] < frame_len:
        return torch.empty((0, frame_len, 2 * fsst.num_rows), device=dev), torch.empty((0, frame_len), dtype=torch.int64, device=dev)
    xd = x.to(dev, non_blocking=True)
    feats = fsst.frames(xd, stride, frame_len)
    labels = frame_batch((y - 1).to(dev, non_blocking=True), stride, frame_len)
    return feats, labels.contiguous()


__all__ = ["load_recording_csv", "recording_to_frames", "frame_signal"]
