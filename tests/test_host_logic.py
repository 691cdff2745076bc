"""Host-side logic of the drop-in layer (no GPU): window preparation, band selection, API surface,
moments, sharding and metric definitions."""
import inspect

import numpy as np
import pytest
import torch

from oracle import fsst_oracle as fo


def test_derivative_window_matches_oracle_dtwin():
    from hss.transforms._window import derivative_window

    for w, fs in ((fo.reference_window(), 1000.0), (np.kaiser(256, 10.0), 2000.0)):
        assert np.abs(derivative_window(w, fs) - fo.dtwin(w, fs)).max() < 1e-12


def test_band_rows_match_reference_mask():
    from hss.transforms._window import band_rows

    assert band_rows(1000, 128, (25, 200)) == (4, 25)          # 22 rows -> 44 features (test_dataset.py:67-69)
    assert band_rows(2000, 128, (50, 400)) == (4, 25)          # config 5
    assert band_rows(1000, 128, (31.25, 195.3125)) == (4, 25)  # inclusive on both ends (synchrosqueeze.py:109)
    assert band_rows(1000, 256, (25, 200)) == fo.band_rows(1000, 256, (25, 200))
    with pytest.raises(ValueError):
        band_rows(1000, 128, (1, 2))


def test_fsst_signature_matches_reference():
    from hss.transforms import FSST

    params = list(inspect.signature(FSST.__init__).parameters)
    assert params == ["self", "fs", "window", "abs", "stack", "truncate_freq", "dtype"]   # synchrosqueeze.py:13-21
    f = FSST(1000, window=fo.reference_window(), truncate_freq=(25, 200), stack=True)
    assert f.num_rows == 22 and f.fs == 1000 and f.stack and not f.abs
    assert torch.allclose(f.frequencies(), torch.arange(4, 26) * 7.8125)
    assert FSST(1000, window=np.kaiser(101, 5.0)).num_rows == 51           # any window length 4 .. 1024 (ssq.fsst takes any window)
    with pytest.raises(ValueError):
        FSST(1000, window=np.ones(2000))


def test_segmenter_signature_state_dict_and_rng_order():
    from hss.model.segmenter import HeartSoundSegmenter
    from oracle import lstm_oracle as lo

    sig = inspect.signature(HeartSoundSegmenter.__init__)
    assert list(sig.parameters) == ["self", "input_size", "batch_size", "hidden_size", "bidirectional", "device", "dtype"]
    assert all(p.kind is inspect.Parameter.KEYWORD_ONLY for n, p in sig.parameters.items() if n != "self")
    torch.manual_seed(68)
    m = HeartSoundSegmenter(input_size=44, batch_size=3)
    params, h0, c0 = lo.reference_params(68, 44, 3, 240)
    assert list(m.state_dict().keys()) == lo.PARAM_NAMES
    assert torch.equal(m.h0, h0) and torch.equal(m.c0, c0)
    for k, v in m.state_dict().items():
        assert torch.equal(v, params[k])
    assert sum(p.numel() for p in m.parameters()) == 1_937_284
    m.eval()
    with pytest.raises(RuntimeError, match="Expected hidden"):
        m(torch.zeros(2, 5, 44))                                      # batch mismatch raises like the reference
    m.train()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(3, 5, 44))                                      # training has no CPU fallback either
    with pytest.raises(RuntimeError, match="eval"):
        m.predict(torch.zeros(3, 5, 44))


def test_moments_api():
    from hss.moments import update_mean, update_variance

    v = np.random.default_rng(2).standard_normal(200)
    m = m2 = 0.0
    for k, x in enumerate(v, 1):
        m2 = update_variance(x, m, m2, k)
        m = update_mean(m, x, k)
        assert update_variance(x, 0.0, 0.0, 1) == 0.0 or k > 1
    assert abs(m - v.mean()) < 1e-12 and abs(m2 / 199 - v.var(ddof=1)) < 1e-12
    assert fo.update_mean(1.0, 3.0, 2) == update_mean(1.0, 3.0, 2)


def test_shard_range_partitions():
    from hss.sharding import shard_range

    for n, w in ((4096, 8), (50, 8), (7, 2), (0, 4), (3, 5)):
        parts = [shard_range(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in parts]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_metrics_from_counts():
    from hss.sharding import metrics_from_counts

    cm = torch.tensor([[8, 2, 0, 0], [1, 9, 0, 0], [0, 0, 5, 5], [0, 0, 0, 10]])
    m = metrics_from_counts(cm)
    assert torch.allclose(m["recall_per_class"], torch.tensor([0.8, 0.9, 0.5, 1.0], dtype=torch.float64))
    assert torch.allclose(m["precision_per_class"], torch.tensor([8 / 9, 9 / 11, 1.0, 10 / 15], dtype=torch.float64))
    assert abs(m["micro_accuracy"] - 32 / 40) < 1e-12
    assert metrics_from_counts(torch.zeros(4, 4))["f1"] == 0.0


def test_macro_average_skips_absent_classes():
    """torchmetrics average="macro" (reference main.py:49-56) gives weight 0 to a class with tp + fp + fn == 0: a shard in which
    class 3 occurs neither in the targets nor in the predictions averages over the three classes that are present."""
    from hss.sharding import metrics_from_counts, metrics_from_state

    cm = torch.tensor([[8, 2, 0, 0], [1, 9, 0, 0], [0, 5, 5, 0], [0, 0, 0, 0]])
    m = metrics_from_counts(cm)
    rec = [0.8, 0.9, 0.5]
    prec = [8 / 9, 9 / 16, 1.0]
    f1 = [2 * p * r / (p + r) for p, r in zip(prec, rec)]
    assert abs(m["recall"] - sum(rec) / 3) < 1e-12 and abs(m["accuracy"] - sum(rec) / 3) < 1e-12
    assert abs(m["precision"] - sum(prec) / 3) < 1e-12 and abs(m["f1"] - sum(f1) / 3) < 1e-12
    # a class that is predicted but never a target IS present (fp > 0) and counts with recall 0
    cm2 = torch.tensor([[8, 2, 0, 1], [1, 9, 0, 0], [0, 5, 5, 0], [0, 0, 0, 0]])
    assert abs(metrics_from_counts(cm2)["recall"] - (8 / 11 + 0.9 + 0.5 + 0.0) / 4) < 1e-12
    # the 18-scalar state: counts + loss sum + count
    st = torch.cat([cm.reshape(-1).double(), torch.tensor([15.0, 30.0], dtype=torch.float64)])
    out = metrics_from_state(st)
    assert out["loss"] == 0.5 and out["count"] == 30 and abs(out["f1"] - m["f1"]) < 1e-15


def test_frame_signal_matches_reference_contract():
    """hss.utils.preprocess: L = floor((T - n) / stride) frames, truncated single frame otherwise (reference preprocess.py:39-56)."""
    import math

    import torch
    from hss.utils.preprocess import frame_batch, frame_signal

    for T, stride, n in ((35000, 1000, 2000), (4001, 1000, 2000), (2000, 1000, 2000), (1500, 1000, 2000), (9999, 333, 512)):
        x = torch.arange(T, dtype=torch.float32)
        y = (torch.arange(T) % 4).to(torch.int64)
        frames, labels = frame_signal(x, y, stride, n)
        L = math.floor((T - n) / stride)
        if L <= 0:
            assert len(frames) == 1 and frames[0].shape == (min(T, n), 1) and torch.equal(frames[0][:, 0], x[:n])
            assert frame_batch(x, stride, n).shape == (1, min(T, n))
            continue
        assert len(frames) == len(labels) == L
        for i, (f, l) in enumerate(zip(frames, labels)):
            assert f.shape == (n, 1) and torch.equal(f[:, 0], x[i * stride:i * stride + n]) and torch.equal(l[:, 0], y[i * stride:i * stride + n])
        fb = frame_batch(x, stride, n)
        assert fb.shape == (L, n) and torch.equal(fb, torch.stack([f[:, 0] for f in frames]))


def test_load_recording_csv_matches_the_reference_reader(tmp_path):
    """hss.utils.ingest.load_recording_csv == pd.read_csv(skiprows=1, names=[Signals, Labels]) (reference heart_sounds.py:193-197)."""
    import numpy as np
    import pandas as pd
    import torch
    from hss.utils.ingest import load_recording_csv
    from workloads import synth_pcg, synthetic_targets

    x = synth_pcg(5000, seed=3).astype(np.float64)
    y = synthetic_targets(1, 5000)[0] + 1                       # file labels are 1..4
    path = tmp_path / "0001.csv"
    with open(path, "w") as f:
        f.write("Signals,Labels\n")
        for a, b in zip(x, y):
            f.write(f"{float(a)!r},{int(b)}\n")
    xs, ys = load_recording_csv(str(path))
    df = pd.read_csv(path, skiprows=1, names=["Signals", "Labels"])
    assert torch.equal(xs, torch.tensor(df.loc[:, "Signals"].to_numpy(), dtype=torch.float32))
    assert torch.equal(ys, torch.tensor(df.loc[:, "Labels"].to_numpy(), dtype=torch.int64))
    assert xs.shape == (5000,) and ys.dtype == torch.int64 and int(ys.min()) == 1 and int(ys.max()) == 4


def test_resample_matches_scipy_fourier_method():
    """hss.transforms.Resample == scipy.signal.resample, the call behind reference hss/transforms/resample.py:21."""
    import numpy as np
    import scipy.signal
    import torch
    from hss.transforms import Resample

    rng = np.random.default_rng(4)
    for n, num in ((2000, 1000), (2000, 4000), (1001, 500), (1000, 1501), (35000, 17500), (64, 64), (7, 12)):
        x = rng.standard_normal(n)
        ref = scipy.signal.resample(x, num)
        got = Resample(num)(torch.from_numpy(x), dtype=torch.float64).numpy()
        assert got.shape == (num,) and np.abs(got - ref).max() < 1e-10, (n, num, np.abs(got - ref).max())
    y = (np.arange(2000) % 4 + 1).astype(np.float64)          # label track, as heart_sounds.py:206 resamples it
    assert Resample(1000)(torch.from_numpy(y)).dtype == torch.float32


def _scored_labels(n, seed, sharp=2.0):
    g = torch.Generator().manual_seed(seed)
    target = torch.randint(0, 4, (n,), generator=g)
    logits = torch.randn(n, 4, generator=g) + sharp * torch.nn.functional.one_hot(target, 4) * (torch.rand(n, 1, generator=g) > 0.3)
    return torch.log_softmax(logits, dim=1), target


def test_auroc_oracle_matches_sklearn():
    """oracle/metrics_oracle.auroc_exact (main.py:48,60 definition) against scikit-learn's independent implementation."""
    from sklearn.metrics import roc_auc_score

    from oracle import metrics_oracle as mo

    logp, target = _scored_labels(5000, 3)
    ours = mo.auroc_exact(logp.numpy(), target.numpy())
    p = np.exp(logp.numpy())
    for c in range(4):
        assert abs(ours[c] - roc_auc_score((target.numpy() == c).astype(int), p[:, c])) < 1e-12
    q = mo.auroc_binned(logp.numpy(), target.numpy(), 64)                       # ties inside a bin share a trapezoid
    qs = np.minimum(63, np.floor(p * np.float32(64)))
    for c in range(4):
        assert abs(q[c] - roc_auc_score((target.numpy() == c).astype(int), qs[:, c])) < 1e-12
    assert mo.auroc_exact(logp.numpy(), np.zeros(5000, np.int64))[1] == 0.0     # class without positives


def test_auroc_from_histograms_host_logic():
    """hss.sharding.auroc_from_histograms on oracle-built histograms == the oracle's AUROC of the binned scores; at 4096
    bins that is the exact AUROC to ~1e-4; shards add up."""
    from hss.sharding import auroc_from_histograms

    from oracle import metrics_oracle as mo

    logp, target = _scored_labels(20000, 5)
    for nbins in (16, 4096):
        h = torch.from_numpy(mo.histograms(logp.numpy(), target.numpy(), nbins))
        got = auroc_from_histograms(h)
        ref = mo.auroc_binned(logp.numpy(), target.numpy(), nbins)
        assert np.abs(got["auroc_per_class"].numpy() - ref).max() < 1e-12
        assert abs(got["auroc"] - ref.mean()) < 1e-12
    exact = mo.auroc_exact(logp.numpy(), target.numpy())
    assert np.abs(got["auroc_per_class"].numpy() - exact).max() < 2e-4
    halves = sum(torch.from_numpy(mo.histograms(logp[s].numpy(), target[s].numpy(), 4096)) for s in (slice(0, 7000), slice(7000, None)))
    assert torch.equal(halves, h)
    empty = auroc_from_histograms(torch.zeros(4, 2, 16, dtype=torch.int64))
    assert empty["auroc"] == 0.0


def test_native_batch_csv_parser(tmp_path, built):
    """hssb_csv_scan / hssb_csv_parse: many files, host thread pool, one staging buffer -- the same values as
    pd.read_csv(skiprows=1) + torch.tensor(dtype) of reference heart_sounds.py:193-197; ragged lengths, CRLF, blank lines, an
    empty recording, exponent notation, a file without trailing newline; errors for missing / malformed files."""
    import numpy as np
    import pandas as pd
    import pytest
    import torch
    from hss.utils import load_recording_csv, load_recordings_csv, scan_recordings_csv

    rng = np.random.default_rng(0)
    paths, lens = [], [1, 7, 35000, 0, 2500, 64]
    for i, n in enumerate(lens):
        x = (rng.standard_normal(n) * 10.0 ** rng.integers(-6, 3, size=n)).astype(np.float64)
        y = rng.integers(1, 5, size=n)
        eol = "\r\n" if i == 1 else "\n"
        path = tmp_path / f"{i:04d}.csv"
        with open(path, "w", newline="") as f:
            f.write("Signals,Labels" + eol)
            for k, (a, b) in enumerate(zip(x, y)):
                txt = f"{a:.17g}" if k % 3 else f"{a:.9e}"
                f.write(f"{txt},{int(b)}" + (eol if (k + 1 < n or i != 4) else ""))      # file 4: no trailing newline
            if i == 5:
                f.write(eol + eol)                                                       # trailing blank lines
        paths.append(str(path))
    rows = scan_recordings_csv(paths, threads=3)
    assert rows.tolist() == lens
    rec = load_recordings_csv(paths, threads=3)
    assert len(rec) == len(paths)
    for i, path in enumerate(paths):
        x, y = rec[i]
        df = pd.read_csv(path, skiprows=1, names=["Signals", "Labels"])
        if lens[i] == 0:
            assert x.numel() == 0 and y.numel() == 0 and len(df) == 0
        else:
            assert torch.equal(x, torch.tensor(df.loc[:, "Signals"].to_numpy(), dtype=torch.float32)), path
            assert torch.equal(y, torch.tensor(df.loc[:, "Labels"].to_numpy(), dtype=torch.int64))
        x1, y1 = load_recording_csv(path)
        assert torch.equal(x1, x) and torch.equal(y1, y) and x1.dtype == torch.float32 and y1.dtype == torch.int64
    assert load_recording_csv(paths[2], dtype=torch.float64)[0].dtype == torch.float64
    with pytest.raises(ValueError):
        load_recordings_csv([paths[0], str(tmp_path / "missing.csv")])
    bad = tmp_path / "bad.csv"
    bad.write_text("Signals,Labels\n0.5,1\nnot-a-number,2\n")
    with pytest.raises(ValueError, match="malformed"):
        load_recordings_csv([str(bad)])
    one = tmp_path / "one_column.csv"
    one.write_text("Signals\n0.5\n")
    with pytest.raises(ValueError):
        load_recordings_csv([str(one)])


def test_recurrent_weight_gradient_on_row_shifted_views():
    """hss/model/_train.py computes dG^T h_prev of a layer without building h_prev: one product on views of the flattened
    outputs shifted by a row plus a B-row correction (h0 at each window's first / last step).  Same result as the product with
    the explicitly concatenated h_prev, for both directions, including T = 1 and B = 1."""
    import torch
    from hss.model._train import edge_fixup, shifted_rows

    g = torch.Generator().manual_seed(0)
    for B, T, H in ((3, 5, 4), (1, 7, 3), (4, 1, 2), (1, 1, 2)):
        M = B * T
        dG = torch.randn(2, M, 4 * H, generator=g, dtype=torch.float64)
        out = torch.randn(B, T, 2 * H, generator=g, dtype=torch.float64)
        h0 = torch.randn(2, B, H, generator=g, dtype=torch.float64)
        hp = (torch.cat([h0[0].unsqueeze(1), out[:, :-1, :H]], dim=1).reshape(M, H),
              torch.cat([out[:, 1:, H:], h0[1].unsqueeze(1)], dim=1).reshape(M, H))
        o2 = out.reshape(M, 2 * H)
        for d in range(2):
            rg, ro, co = shifted_rows(M, H, d)
            edge_rows = torch.arange(B) * T + (0 if d == 0 else T - 1)
            got = dG[d][rg].t() @ o2[ro, co] + edge_fixup(dG[d][edge_rows], o2, h0[d], B, T, H, d)
            assert torch.allclose(got, dG[d].t() @ hp[d], rtol=1e-12, atol=1e-12), (B, T, d)
