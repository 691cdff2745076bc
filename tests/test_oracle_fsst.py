"""Known-answer / property tests that pin the float64 FSST restatement (oracle/fsst_oracle.py).

The reference holds no value-level vector for ssq.fsst (only the (2000, 44) shape,
reference test/test_dataset.py:67-69), so the oracle is pinned analytically.
"""
import ctypes
import os

import numpy as np
import pytest

from oracle import fsst_oracle as fo

FS = 1000.0
W = fo.reference_window()


def test_window_matches_scipy_kaiser():
    import scipy.signal

    ref = scipy.signal.get_window(("kaiser", 0.5), 128, fftbins=False)   # reference main.py:155
    assert np.abs(ref - W).max() < 1e-15


def test_dtwin_matches_scipy_notaknot_spline():
    from scipy.interpolate import CubicSpline

    for w in (W, np.kaiser(256, 10.0), np.hanning(65)):
        n = len(w)
        ref = CubicSpline(np.arange(1, n + 1), w, bc_type="not-a-knot")(np.arange(1, n + 1), 1)
        assert np.abs(fo.notaknot_slopes(w) - ref).max() < 1e-13
    dg = fo.dtwin(W, FS)
    assert abs(dg[0] - 0.2946) < 1e-4 and abs(dg[63] - 0.00239) < 1e-5   # SURVEY 8c probe values
    assert np.abs(dg + dg[::-1]).max() < 1e-12                           # antisymmetric


def test_shape_and_band_match_reference_test():
    # reference test/test_dataset.py:67-69: x.shape == (2000, 44)
    x = fo.synth_pcg(2000)
    feats = fo.fsst_features(x, FS, W, stack=True, truncate_freq=(25, 200))
    assert feats.shape == (2000, 44) and feats.dtype == np.float32
    assert fo.band_rows(FS, 128, (25, 200)) == (4, 25)
    assert fo.band_rows(2000.0, 128, (50, 400)) == (4, 25)
    s, f, t = fo.fsst(x, FS, W)
    assert s.shape == (65, 2000) and f.shape == (65,) and t.shape == (2000,)
    assert f[1] == 7.8125 and t[1] == 1e-3


def test_zero_input_is_zero_without_nan_and_stack_gives_nan():
    s, _, _ = fo.fsst(np.zeros(500), FS, W)
    assert not np.isnan(s).any() and np.abs(s).max() == 0.0
    feats = fo.fsst_features(np.zeros(500), FS, W, stack=True, truncate_freq=(25, 200))
    assert np.isnan(feats).all()      # 0/0 in the reference's z-score (synchrosqueeze.py:84-85)


@pytest.mark.parametrize("k0", [8, 16])
def test_bin_centred_cosine_is_squeezed_into_its_row(k0):
    t = np.arange(2000) / FS
    s, _, _ = fo.fsst(np.cos(2 * np.pi * k0 * 7.8125 * t), FS, W)
    col = np.abs(s[:, 1000])
    assert col.argmax() == k0
    assert (col[k0] ** 2) / (col ** 2).sum() > 0.99999
    assert abs(col[k0] - 128 * W[64] / 2) < 0.05        # ~63.97 / 64.01


def test_reconstruction_identity_and_column_conservation():
    x = fo.synth_pcg(1500, seed=3).astype(np.float64)
    s, _, _, rows = fo.fsst(x, FS, W, return_rows=True)
    rec = 2 * np.real(s.sum(0)) - np.real(s[0]) - np.real(s[64])
    assert np.abs(rec - 128 * W[64] * x).max() < 1e-10
    assert rows.min() >= 0 and rows.max() <= 127


def test_stacked_features_are_standardised_and_scale_invariant():
    x = fo.synth_pcg(2000, seed=11)
    a = fo.fsst_features(x, FS, W, stack=True, truncate_freq=(25, 200))
    for half in (a[:, :22], a[:, 22:]):
        assert abs(half.mean()) < 1e-6 and abs(half.std(ddof=1) - 1.0) < 1e-5
    b = fo.fsst_features(3.7 * x.astype(np.float64), FS, W, stack=True, truncate_freq=(25, 200))
    assert np.abs(a - b).max() < 5e-6


def test_chirp_ridge_follows_instantaneous_frequency_kaiser256():
    # MATLAB's default window kaiser(256, 10) exercises multi-bin reassignment
    w = np.kaiser(256, 10.0)
    n = 4000
    t = np.arange(n) / FS
    f0, f1 = 50.0, 250.0
    x = np.cos(2 * np.pi * (f0 * t + 0.5 * (f1 - f0) / t[-1] * t ** 2))
    s, f, _ = fo.fsst(x, FS, w)
    for ti in (1000, 2000, 3000):
        inst = f0 + (f1 - f0) * t[ti] / t[-1]
        assert abs(f[np.abs(s[:, ti]).argmax()] - inst) <= FS / 256


def test_moments_recurrences_match_numpy():
    rng = np.random.default_rng(0)
    v = rng.standard_normal(500)
    m, m2 = 0.0, 0.0
    for k, x in enumerate(v, start=1):
        m2 = fo.update_variance(x, m, m2, k)
        m = fo.update_mean(m, x, k)
    assert abs(m - v.mean()) < 1e-12 and abs(m2 / (len(v) - 1) - v.var(ddof=1)) < 1e-12


def test_c_restatement_matches_numpy_restatement(built):
    lib = ctypes.CDLL(os.path.join(os.path.dirname(fo.__file__), "_build", "libhss_oracle.so"))
    P = ctypes.c_void_p
    lib.hsso_fsst.argtypes = [P, ctypes.c_int64, ctypes.c_int64, ctypes.c_double, P, P, ctypes.c_int, P]
    lib.hsso_fsst_features.argtypes = [P, ctypes.c_int64, ctypes.c_int64, ctypes.c_double, P, P, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, P]
    B, N = 3, 1111
    x = fo.synth_pcg_batch(B, N, seed=21)
    dg = fo.dtwin(W, FS)
    s = np.zeros((B, 65, N), dtype=np.complex128)
    assert lib.hsso_fsst(x.ctypes.data, B, N, FS, W.ctypes.data, dg.ctypes.data, 128, s.ctypes.data) == 0
    out = np.zeros((B, N, 44), dtype=np.float32)
    assert lib.hsso_fsst_features(x.ctypes.data, B, N, FS, W.ctypes.data, dg.ctypes.data, 128, 4, 25, 2, out.ctypes.data) == 0
    for b in range(B):
        ref, _, _ = fo.fsst(x[b], FS, W)
        assert np.abs(ref - s[b]).max() < 1e-12 * np.abs(ref).max()
        feats = fo.fsst_features(x[b], FS, W, stack=True, truncate_freq=(25, 200))
        assert np.abs(feats - out[b]).max() < 2e-6


def test_golden_vectors_regress(golden_dir):
    for name in ("fsst_pcg.npz", "fsst_noise.npz"):
        g = np.load(os.path.join(golden_dir, name))
        feats = fo.fsst_features(g["x"], float(g["fs"]), g["window"], stack=True, truncate_freq=tuple(g["truncate"]))
        assert np.array_equal(feats, g["features"])
        s, _, _ = fo.fsst(g["x"], float(g["fs"]), g["window"])
        k_lo, k_hi = g["band"]
        assert np.array_equal(s[k_lo:k_hi + 1].astype(np.complex64), g["s_band"])
