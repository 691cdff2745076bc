"""Synthetic PCG workloads shared by the tests, ``smoke()`` and ``bench.py``.

The real dataset (David Springer heart sounds) is a network download (reference
hss/datasets/heart_sounds.py:136-151) and is not available offline, so every configuration of
BASELINE.json runs on these seeded generators.
"""
from __future__ import annotations

import numpy as np


def synth_pcg(n: int, fs: float = 1000.0, seed: int = 68) -> np.ndarray:
    """0.05*N(0,1) noise + S1 (55 Hz) / S2 (90 Hz) Gaussian bursts every 0.83 s.  float32 [n]."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    x = 0.05 * rng.standard_normal(n)
    c = rng.uniform(0.0, 0.83)
    while c < t[-1] + 0.5:
        x += np.exp(-((t - c) / 0.03) ** 2) * np.sin(2 * np.pi * 55.0 * (t - c))
        x += 0.7 * np.exp(-((t - c - 0.3) / 0.02) ** 2) * np.sin(2 * np.pi * 90.0 * (t - c - 0.3))
        c += 0.83
    return x.astype(np.float32)


def synth_pcg_batch(b: int, n: int, fs: float = 1000.0, seed: int = 68) -> np.ndarray:
    return np.stack([synth_pcg(n, fs, seed + i) for i in range(b)]) if b else np.zeros((0, n), np.float32)


def tiled_windows(n_windows: int, n: int, fs: float = 1000.0, seed: int = 68) -> np.ndarray:
    """``n_windows`` windows built from 32 distinct synthetic recordings with per-window gains."""
    base = synth_pcg_batch(32, n, fs, seed)
    reps = (n_windows + 31) // 32
    gains = np.linspace(0.5, 2.0, reps * 32, dtype=np.float32)[:, None]
    return (np.tile(base, (reps, 1)) * gains)[:n_windows].copy()


def synthetic_targets(n_windows: int, n: int) -> np.ndarray:
    """Cyclic S1/systole/S2/diastole labels 0..3 with durations 120/200/100/410 ms (int64)."""
    pattern = np.concatenate([np.full(d, c, dtype=np.int64) for c, d in enumerate((120, 200, 100, 410))])
    row = np.tile(pattern, n // len(pattern) + 1)[:n]
    return np.tile(row, (n_windows, 1))


def reference_window(nwin: int = 128, beta: float = 0.5) -> np.ndarray:
    """``scipy.signal.get_window(("kaiser", 0.5), 128, fftbins=False)`` (reference main.py:155);
    np.kaiser is the same symmetric Kaiser window."""
    return np.kaiser(nwin, beta)
