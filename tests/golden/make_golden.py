"""Generates the committed golden fixtures.  Run in the BUILD container only:

    python tests/golden/make_golden.py

* ``lstm_*.npz``: outputs of the REAL reference ``HeartSoundSegmenter`` (imported by file path from
  /root/reference/hss/model/segmenter.py -- that module only needs torch) on seeded inputs.
* ``fsst_*.npz``: the reference's FSST (``ssq 0.1.0``) is not obtainable (see oracle/fsst_oracle.py),
  so these vectors come from the float64 numpy restatement and pin *it* (and, through the parity
  tests, the CUDA kernels); they are regression vectors, not libssq outputs.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import fsst_oracle as fo  # noqa: E402

REF = "/root/reference/hss/model/segmenter.py"


def load_reference_segmenter():
    spec = importlib.util.spec_from_file_location("ref_segmenter", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.HeartSoundSegmenter


def lstm_case(name, seed, B, T, F, H, store_params):
    Seg = load_reference_segmenter()
    torch.manual_seed(seed)
    model = Seg(input_size=F, batch_size=B, hidden_size=H)
    model.eval()
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, T, F, generator=g)
    with torch.no_grad():
        logp = model(x)
    out = {"seed": seed, "B": B, "T": T, "F": F, "H": H, "x": x.numpy(), "logp": logp.numpy(),
           "h0": model.h0.numpy(), "c0": model.c0.numpy()}
    sd = model.state_dict()
    if store_params:
        for k, v in sd.items():
            out["param:" + k] = v.numpy()
    else:  # params are regenerated from the seed by oracle.lstm_oracle.reference_params; pin a checksum
        out["param_sum"] = np.array([float(v.double().sum()) for v in sd.values()])
    out["param_names"] = np.array(list(sd.keys()))
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "logp", logp.shape)


def fsst_case(name, x, fs, window, truncate):
    s, f, t = fo.fsst(x, fs, window)
    k_lo, k_hi = fo.band_rows(fs, len(window), truncate)
    feats = fo.fsst_features(x, fs, window, stack=True, truncate_freq=truncate)
    mags = fo.fsst_features(x, fs, window, abs=True, truncate_freq=truncate)
    np.savez_compressed(os.path.join(HERE, name), x=x.astype(np.float32), fs=fs, window=window,
                        band=np.array([k_lo, k_hi]), s_band=s[k_lo:k_hi + 1].astype(np.complex64),
                        features=feats, magnitudes=mags, truncate=np.array(truncate, dtype=np.float64))
    print(name, "features", feats.shape)


if __name__ == "__main__":
    lstm_case("lstm_small.npz", seed=7, B=3, T=40, F=44, H=16, store_params=True)
    lstm_case("lstm_h240.npz", seed=68, B=2, T=64, F=44, H=240, store_params=False)
    w = fo.reference_window()
    fsst_case("fsst_pcg.npz", fo.synth_pcg(2000, 1000.0, 68), 1000.0, w, (25, 200))
    rng = np.random.default_rng(5)
    fsst_case("fsst_noise.npz", rng.standard_normal(777).astype(np.float32), 1000.0, w, (25, 200))
