"""Recording ingest for the B200 path: CSV files -> pinned host staging -> async H2D -> framed, transformed batch.

Reference: ``hss/datasets/heart_sounds.py:155-169,193-197``.  ``_load_file`` reads one two-column CSV (header row, column 0 the
PCG signal, column 1 the labels 1..4) with pandas, the constructor then frames it and calls the transform once per frame.  Here

* ``load_recordings_csv`` parses MANY files with the native batch parser of ``libhssb.so`` (``hssb_csv_scan`` /
  ``hssb_csv_parse``: a pool of host threads, decimal -> float64 -> float32 like the reference) straight into one pinned
  staging buffer;
* ``load_recording_csv`` is the single-file form with ``_load_file``'s contract;
* ``recording_to_frames`` is the in-memory + framing branch of the dataset constructor as ONE device call: frames of
  ``frame_len`` samples every ``stride`` samples, labels shifted to 0..3 (``y - 1``, ``:164``), every frame transformed by
  ``FSST`` (``:166``) -- recordings shorter than one frame are skipped (``:160-161``);
* ``stream_recordings`` pipelines a whole file list: while the GPU transforms the recordings of one group of files, a host
  thread parses the next group into the other staging buffer and its H2D copies are issued on a copy stream.
"""
from __future__ import annotations

import ctypes
import threading
from typing import Iterable, Iterator, Sequence

import numpy as np
import torch

from .. import _lib
from .preprocess import frame_batch, frame_signal


def _paths_array(paths: Sequence[str]):
    arr = (ctypes.c_char_p * len(paths))()
    arr[:] = [str(p).encode() for p in paths]
    return arr


def scan_recordings_csv(paths: Sequence[str], threads: int = 0) -> np.ndarray:
    """Number of samples (data rows) of every file."""
    rows = np.zeros(len(paths), dtype=np.int64)
    if len(paths):
        rc = _lib.lib().hssb_csv_scan(_paths_array(paths), len(paths), threads, rows.ctypes.data)
        _lib.check(rc, "hssb_csv_scan")
    return rows


class ParsedRecordings:
    """Signals / labels of a group of files in one (pinned) staging buffer; ``self[i] -> (x_i, y_i)`` views."""

    def __init__(self, signal: torch.Tensor, labels: torch.Tensor, offsets: np.ndarray):
        self.signal, self.labels, self.offsets = signal, labels, offsets

    def __len__(self) -> int:
        return len(self.offsets) - 1

    def __getitem__(self, i: int) -> tuple[torch.Tensor, torch.Tensor]:
        lo, hi = int(self.offsets[i]), int(self.offsets[i + 1])
        return self.signal[lo:hi], self.labels[lo:hi]


def _staging(n: int, pin: bool) -> tuple[torch.Tensor, torch.Tensor]:
    x = torch.empty(max(n, 1), dtype=torch.float32)
    y = torch.empty(max(n, 1), dtype=torch.int64)
    if pin and torch.cuda.is_available():
        x, y = x.pin_memory(), y.pin_memory()
    return x, y


def load_recordings_csv(paths: Sequence[str], pin: bool = False, threads: int = 0,
                        staging: tuple[torch.Tensor, torch.Tensor] | None = None) -> ParsedRecordings:
    """Parse all files (host thread pool) into one staging buffer; ``staging`` reuses a previous (large enough) one."""
    paths = list(paths)
    rows = scan_recordings_csv(paths, threads)
    offsets = np.zeros(len(paths) + 1, dtype=np.int64)
    np.cumsum(rows, out=offsets[1:])
    total = int(offsets[-1])
    if staging is None or staging[0].numel() < total:
        staging = _staging(total, pin)
    x, y = staging
    if paths:
        rc = _lib.lib().hssb_csv_parse(_paths_array(paths), len(paths), threads, offsets.ctypes.data, x.data_ptr(), y.data_ptr())
        _lib.check(rc, "hssb_csv_parse")
    return ParsedRecordings(x, y, offsets)


def load_recording_csv(path: str, dtype: torch.dtype = torch.float32, pin: bool = False) -> tuple[torch.Tensor, torch.Tensor]:
    """``(signal [T] dtype, labels [T] int64)`` of one recording file (header row skipped), optionally in pinned memory."""
    rec = load_recordings_csv([path], pin=pin, threads=1)
    x, y = rec[0]
    if x.numel() == 0:
        return x.to(dtype), y
    return (x if dtype == torch.float32 else x.to(dtype)), y


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


def stream_recordings(paths: Iterable[str], fsst, stride: int = 1000, frame_len: int = 2000, device: torch.device | str | None = None,
                      group: int = 16, threads: int = 0) -> Iterator[tuple[torch.Tensor, torch.Tensor]]:
    """Yield ``(features [L, frame_len, F], labels [L, frame_len])`` for every recording of a file list (short ones skipped).

    Three-stage pipeline over groups of ``group`` files: a host thread parses group k+1 into the free pinned staging buffer
    (``hssb_csv_parse`` releases the GIL), the H2D copies of group k run on a copy stream, and the FSST of every recording runs on
    the caller's current stream behind its copy -- the transform of one recording overlaps the ingest of the next ones.
    """
    paths = list(paths)
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    groups = [paths[i:i + group] for i in range(0, len(paths), group)]
    if not groups:
        return
    copy_stream = torch.cuda.Stream(dev)
    staging: list = [None, None]
    released = [None, None]                 # event after the last H2D copy out of a staging buffer
    parsed: dict[int, ParsedRecordings] = {}
    errors: list[BaseException] = []

    def parse(k: int):
        try:
            if released[k & 1] is not None:
                released[k & 1].synchronize()
            parsed[k] = load_recordings_csv(groups[k], pin=True, threads=threads, staging=staging[k & 1])
            staging[k & 1] = (parsed[k].signal, parsed[k].labels)
        except BaseException as e:          # surfaced on the consumer thread
            errors.append(e)

    worker = threading.Thread(target=parse, args=(0,))
    worker.start()
    for k in range(len(groups)):
        worker.join()
        if errors:
            raise errors[0]
        rec = parsed.pop(k)
        if k + 1 < len(groups):
            worker = threading.Thread(target=parse, args=(k + 1,))
            worker.start()
        with torch.cuda.device(dev):
            compute = torch.cuda.current_stream(dev)
            staged = []
            with torch.cuda.stream(copy_stream):
                total = int(rec.offsets[-1])
                xd = rec.signal[:total].to(dev, non_blocking=True)
                yd = (rec.labels[:total].to(dev, non_blocking=True) - 1)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                released[k & 1] = ev
            compute.wait_event(ev)
            xd.record_stream(compute)
            yd.record_stream(compute)
            for i in range(len(rec)):
                lo, hi = int(rec.offsets[i]), int(rec.offsets[i + 1])
                if hi - lo < frame_len:
                    continue                # heart_sounds.py:160-161
                feats = fsst.frames(xd[lo:hi], stride, frame_len)
                labels = frame_batch(yd[lo:hi], stride, frame_len).contiguous()
                staged.append((feats, labels))
            yield from staged
    if worker.is_alive():
        worker.join()


__all__ = ["load_recording_csv", "load_recordings_csv", "scan_recordings_csv", "recording_to_frames", "stream_recordings",
           "frame_signal", "ParsedRecordings"]
