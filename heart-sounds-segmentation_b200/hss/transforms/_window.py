"""Host-side window preparation for the FSST kernels (float64, once per ``FSST`` instance).

``ssq.fsst`` (reference hss/transforms/synchrosqueeze.py:48) derives the "derivative window" from
the analysis window inside the library (MATLAB ``dtwin``): the analytic derivative of the
not-a-knot cubic spline through the window samples, evaluated at the samples, times fs/(2*pi).
Here it is a banded linear solve for the knot slopes.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import solve_banded


def derivative_window(window: np.ndarray, fs: float) -> np.ndarray:
    g = np.asarray(window, dtype=np.float64).reshape(-1)
    n = g.size
    if n < 4:
        raise ValueError("FSST window needs at least 4 samples")
    step = np.diff(g)
    # unknowns: slopes s[0..n-1] on unit-spaced knots.
    # C2 continuity at interior knots:  s[i-1] + 4 s[i] + s[i+1] = 3 (step[i-1] + step[i])
    # not-a-knot ends (C3 across the 2nd / 2nd-last knot), already reduced to two unknowns each:
    #   s[0] + 2 s[1] = (5 step[0] + step[1]) / 2 ;  2 s[n-2] + s[n-1] = (step[n-3] + 5 step[n-2]) / 2
    ab = np.zeros((3, n))
    ab[0, 1:] = 1.0       # super-diagonal
    ab[1, :] = 4.0        # diagonal
    ab[2, :-1] = 1.0      # sub-diagonal
    ab[1, 0], ab[0, 1] = 1.0, 2.0
    ab[1, -1], ab[2, -2] = 1.0, 2.0
    rhs = np.empty(n)
    rhs[1:-1] = 3.0 * (step[:-1] + step[1:])
    rhs[0] = 0.5 * (5.0 * step[0] + step[1])
    rhs[-1] = 0.5 * (step[-2] + 5.0 * step[-1])
    slopes = solve_banded((1, 1), ab, rhs)
    return slopes * (float(fs) / (2.0 * np.pi))


def band_rows(fs: float, nfft: int, truncate_freq) -> tuple[int, int]:
    """Inclusive bin range kept by the reference band mask (synchrosqueeze.py:107-111).

    The reference compares a float32 frequency tensor with the two Python scalars using
    ``>=`` / ``<=``; bins are ascending so the mask is one contiguous range.
    """
    import torch

    k = nfft // 2 + 1
    f = torch.tensor(np.arange(k) * (float(fs) / nfft), dtype=torch.float32)
    lo, hi = truncate_freq
    keep = torch.nonzero(torch.logical_and(f >= lo, f <= hi)).reshape(-1)
    if keep.numel() == 0:
        raise ValueError(f"truncate_freq={truncate_freq} keeps no frequency bin (fs={fs}, nfft={nfft})")
    return int(keep[0]), int(keep[-1])
