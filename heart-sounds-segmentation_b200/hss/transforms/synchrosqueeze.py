"""``hss.transforms.FSST`` backed by the sm_100a kernels of ``libhssb.so``.

Same constructor, call signature, output shapes / dtypes and branch precedence
(truncate -> abs -> stack -> raw) as reference ``hss/transforms/synchrosqueeze.py:8-111``;
``ssq.fsst`` (synchrosqueeze.py:48) and the ten torch post-processing ops
(synchrosqueeze.py:50-89) are replaced by three CUDA kernels behind ``hssb_fsst_forward``.

Additions that do not change the per-item semantics:
  * ``FSST.batch(x[B, N])`` transforms a batch of equal-length windows in one call;
  * CUDA tensors are accepted (CUDA in -> CUDA out); CPU tensors are copied to the current CUDA
    device and the result is copied back (CPU in -> CPU out), as the dataset code expects
    (reference hss/datasets/heart_sounds.py:199-201).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import numpy.typing as npt
import torch

from .. import _lib
from ._window import band_rows, derivative_window


class FSST:
    """Fourier Synchrosqueezed Transform (reference synchrosqueeze.py:8-35)."""

    def __init__(
        self,
        fs: float,
        window: npt.NDArray,
        abs: bool = False,
        stack: bool = False,
        truncate_freq: Optional[tuple] = None,
        dtype: torch.dtype = torch.float32,
    ):
        self.fs: float = fs
        self.window: npt.NDArray = window
        self.abs = abs
        self.stack = stack
        self.truncate_freq = truncate_freq
        self.dtype = dtype

        w = np.asarray(window, dtype=np.float64).reshape(-1)
        self._nwin = int(w.size)
        if not 4 <= self._nwin <= 1024:
            raise ValueError(f"FSST (B200 build) supports window lengths 4 .. 1024, got {self._nwin}")
        dw = derivative_window(w, fs)
        self._host_windows = torch.from_numpy(np.concatenate([w, dw]).astype(np.float32))
        self._dev_windows: dict[int, torch.Tensor] = {}
        self._workspace: dict = {}
        if truncate_freq:
            self._k_lo, self._k_hi = band_rows(fs, self._nwin, truncate_freq)
        else:
            self._k_lo, self._k_hi = 0, self._nwin // 2
        self._mode = _lib.MODE_ABS if abs else (_lib.MODE_STACK if stack else _lib.MODE_RAW)

    # ------------------------------------------------------------------------------------------
    @property
    def num_rows(self) -> int:
        return self._k_hi - self._k_lo + 1

    def frequencies(self) -> torch.Tensor:
        """Centre frequencies of the kept rows (the ``f`` the reference computes and discards)."""
        k = torch.arange(self._k_lo, self._k_hi + 1, dtype=torch.float64)
        return (k * (float(self.fs) / self._nwin)).to(self.dtype)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """One signal ``[N]`` or ``[N, 1]`` -> same result as reference synchrosqueeze.py:37-65."""
        if x.dim() == 2 and x.shape[1] == 1:
            x = x[:, 0]
        if x.dim() != 1:
            raise ValueError(f"FSST expects a 1-D signal or [N, 1], got shape {tuple(x.shape)}")
        return self.batch(x.unsqueeze(0))[0]

    def frames(self, x: torch.Tensor, stride: int, n: int) -> torch.Tensor:
        """Transform every ``n``-sample frame (one every ``stride`` samples) of a whole recording in one call.

        Same frames as reference ``hss/utils/preprocess.py:39-56`` and same per-frame result as calling the
        transform on each of them (every frame is zero-padded and z-scored on its own, reference
        ``hss/datasets/heart_sounds.py:160-169``): ``[L, ...]`` stacked along a new first dimension.
        """
        from ..utils.preprocess import frame_batch

        return self.batch(frame_batch(x, stride, n))

    def batch(self, x: torch.Tensor) -> torch.Tensor:
        """``x[B, N]`` -> raw ``[B, Kt, N]`` complex64 | abs ``[B, N, Kt]`` | stack ``[B, N, 2*Kt]`` float32."""
        if x.dim() != 2:
            raise ValueError(f"FSST.batch expects [B, N], got shape {tuple(x.shape)}")
        lib = _lib.lib()
        was_cpu = not x.is_cuda
        dev = _lib.require_cuda() if was_cpu else x.device
        with torch.cuda.device(dev):
            xd = x.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
            B, N = xd.shape
            kt = self.num_rows
            if self._mode == _lib.MODE_RAW:
                out = torch.empty((B, kt, N), dtype=torch.complex64, device=dev)
            elif self._mode == _lib.MODE_ABS:
                out = torch.empty((B, N, kt), dtype=torch.float32, device=dev)
            else:
                out = torch.empty((B, N, 2 * kt), dtype=torch.float32, device=dev)
            if B == 0 or N == 0:
                return out.cpu() if was_cpu else out
            win = self._dev_windows.get(dev.index)
            if win is None:
                win = self._host_windows.to(dev)
                self._dev_windows[dev.index] = win
            need = lib.hssb_fsst_workspace_bytes(B, N, self._nwin, self._k_lo, self._k_hi, self._mode)
            ws = _lib.cached_workspace(self._workspace, dev, need)
            rc = lib.hssb_fsst_forward(
                xd.data_ptr(), B, N, win.data_ptr(), win.data_ptr() + 4 * self._nwin, self._nwin,
                float(self.fs), self._k_lo, self._k_hi, self._mode, out.data_ptr(),
                ws.data_ptr(), ws.numel(), _lib.stream_ptr(),
            )
            _lib.check(rc, "hssb_fsst_forward")
        return out.cpu() if was_cpu else out
