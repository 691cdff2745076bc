"""CPU replay of the CUDA kernels' per-thread phases (csrc/fsst_phases.cuh) against the float64
oracle: checks the FFT factorisation, the shared-memory index maps and the reassignment rule
without a GPU."""
import ctypes
import os

import numpy as np
import pytest

from oracle import fsst_oracle as fo


@pytest.fixture(scope="module")
def sim(built):
    lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_sim", "libhssb_sim.so"))
    P = ctypes.c_void_p
    lib.hssb_sim_fsst.argtypes = [P, ctypes.c_longlong, P, P, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int, P, P, P]
    lib.hssb_sim_merge.argtypes = [P, P, P]
    return lib


def run(sim, x, fs, w, k_lo, k_hi):
    n, nwin = len(x), len(w)
    k = nwin // 2 + 1
    g = w.astype(np.float32)
    dg = fo.dtwin(w, fs).astype(np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    sg = np.zeros((k, n), np.complex64)
    sdg = np.zeros((k, n), np.complex64)
    t = np.zeros((k_hi - k_lo + 1, n), np.complex64)
    assert sim.hssb_sim_fsst(x.ctypes.data, n, g.ctypes.data, dg.ctypes.data, nwin, fs, k_lo, k_hi,
                             sg.ctypes.data, sdg.ctypes.data, t.ctypes.data) == 0
    return sg, sdg, t


@pytest.mark.parametrize("nwin,beta,n", [(128, 0.5, 2000), (128, 0.5, 37), (128, 0.5, 1987), (256, 10.0, 700)])
def test_stft_phases_match_fft(sim, nwin, beta, n):
    w = np.kaiser(nwin, beta)
    x = fo.synth_pcg(n, seed=9)
    k = nwin // 2 + 1
    sg, sdg, _ = run(sim, x, 1000.0, w, 0, k - 1)
    rg, rdg = fo.stft_pair(x, 1000.0, w)
    assert np.abs(sg - rg[:k]).max() < 1e-6 * np.abs(rg).max()
    assert np.abs(sdg - rdg[:k]).max() < 2e-6 * np.abs(rdg).max()


def test_reassignment_matches_oracle_rows_and_values(sim):
    w = fo.reference_window()
    x = fo.synth_pcg(2000)
    _, _, t_full = run(sim, x, 1000.0, w, 0, 64)
    s, _, _ = fo.fsst(x, 1000.0, w)
    d = np.abs(t_full - s)
    scale = np.abs(s).max()
    assert (d > 1e-5 * scale).sum() == 0          # no cell landed in a different row than in float64
    assert d.max() < 5e-7 * scale
    _, _, t_band = run(sim, x, 1000.0, w, 4, 25)
    assert np.array_equal(t_band, t_full[4:26])   # band-fused output == slice of the full one


def test_chan_merge_is_welford(sim):
    rng = np.random.default_rng(1)
    a, b = rng.standard_normal(17), rng.standard_normal(40) + 3

    def mom(v):
        return np.array([len(v), v.mean(), ((v - v.mean()) ** 2).sum()])

    out = np.zeros(3)
    ma, mb = mom(a), mom(b)
    sim.hssb_sim_merge(ma.ctypes.data, mb.ctypes.data, out.ctypes.data)
    assert np.allclose(out, mom(np.concatenate([a, b])), rtol=1e-12)
    z = np.zeros(3)
    sim.hssb_sim_merge(z.ctypes.data, z.ctypes.data, out.ctypes.data)
    assert np.array_equal(out, z)
