"""``hss.transforms.Resample`` with the contract of reference ``hss/transforms/resample.py:5-21``.

The reference calls ``scipy.signal.resample`` (Fourier method) on the CPU.  Here the same Fourier
resampling -- keep the common part of the one-sided spectrum, split / merge the Nyquist bin when
the shorter length is even, inverse transform at the new length, scale by ``num / N`` -- runs with
``torch.fft`` on whatever device the signal lives on (CUDA in -> CUDA out, CPU in -> CPU out), so a
``Compose([Resample(n), FSST(...)])`` pipeline never leaves the GPU.  Not on the north-star path
(reference ``main.py:151-160`` does not use it); provided so that ``hss.transforms`` is a complete
drop-in for ``hss/datasets/heart_sounds.py:202-207``.
"""
from __future__ import annotations

import torch


class Resample:
    def __init__(self, num: int) -> None:
        """``num``: number of output samples (reference resample.py:6-11)."""
        self.num = int(num)

    def __call__(self, x: torch.Tensor, dtype: torch.dtype = torch.float32) -> torch.Tensor:
        """Resample ``x`` along its first dimension to ``num`` samples (reference resample.py:13-21)."""
        n, num = int(x.shape[0]), self.num
        if n == 0 or num <= 0:
            raise ValueError(f"cannot resample {n} samples to {num}")
        xf = torch.fft.rfft(x.to(torch.float64), dim=0)
        out = torch.zeros((num // 2 + 1,) + tuple(x.shape[1:]), dtype=xf.dtype, device=x.device)
        nmin = min(n, num)
        nyq = nmin // 2 + 1
        out[:nyq] = xf[:nyq]
        if nmin % 2 == 0:
            if num < n:        # down-sampling: the new Nyquist bin collects the +/- pair of the old spectrum
                out[nmin // 2] = out[nmin // 2] * 2.0
            elif n < num:      # up-sampling: the old Nyquist bin is split between +/- frequencies
                out[nmin // 2] = out[nmin // 2] * 0.5
        y = torch.fft.irfft(out, n=num, dim=0) * (float(num) / float(n))
        return y.to(dtype)
