"""Property tests (hypothesis) of the host logic around the hot path: sharding, framing, metric finalisation.
CPU only; the invariants are the ones the N-GPU job relies on (shards partition the windows, counters add up,
metrics stay in range whatever the counts)."""
import math

import numpy as np
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import metrics_oracle as mo


@settings(max_examples=60, deadline=None)
@given(n=st.integers(0, 10_000), world=st.integers(1, 16))
def test_shards_partition_the_windows(n, world):
    from hss.sharding import shard_range

    parts = [shard_range(n, r, world) for r in range(world)]
    assert parts[0][0] == 0 and parts[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    sizes = [hi - lo for lo, hi in parts]
    assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n


@settings(max_examples=40, deadline=None)
@given(T=st.integers(1, 6000), stride=st.integers(1, 1500), n=st.integers(1, 2500))
def test_framing_covers_the_reference_frame_count(T, stride, n):
    """frame_batch == stacked frame_signal frames; L = floor((T - n) / stride) (reference preprocess.py:39-56)."""
    from hss.utils.preprocess import frame_batch, frame_signal

    x = torch.arange(T, dtype=torch.float32)
    frames, labels = frame_signal(x, torch.zeros(T, dtype=torch.int64), stride, n)
    L = math.floor((T - n) / stride)
    fb = frame_batch(x, stride, n)
    if L <= 0:
        assert len(frames) == 1 and frames[0].shape[0] == min(T, n) and fb.shape == (1, min(T, n))
    else:
        assert len(frames) == len(labels) == L and fb.shape == (L, n)
        assert torch.equal(fb[:, 0], torch.arange(L, dtype=torch.float32) * stride)      # frame i starts at i * stride
        assert torch.equal(fb[-1], frames[-1][:, 0])


@settings(max_examples=50, deadline=None)
@given(counts=st.lists(st.integers(0, 10_000), min_size=16, max_size=16), split=st.integers(0, 16))
def test_metrics_from_counts_bounded_and_additive(counts, split):
    from hss.sharding import metrics_from_counts

    cm = torch.tensor(counts, dtype=torch.int64).reshape(4, 4)
    m = metrics_from_counts(cm)
    ref = mo.per_class(cm.numpy())
    for k in ("recall", "precision", "f1"):
        v = m[f"{k}_per_class"].numpy()
        assert np.all((v >= 0) & (v <= 1)) and np.allclose(v, ref[k], atol=1e-12)
        present = (cm.sum(1) + cm.sum(0)).numpy() > 0           # macro average over the classes that occur (torchmetrics)
        assert abs(m[k] - (ref[k][present].mean() if present.any() else 0.0)) < 1e-12
    # counters of two shards add up to the counters of the whole
    a = cm.clone().reshape(-1)
    a[split:] = 0
    b = cm.reshape(-1) - a
    whole = metrics_from_counts((a + b).reshape(4, 4))
    assert whole["f1"] == m["f1"] and whole["micro_accuracy"] == m["micro_accuracy"]


@settings(max_examples=30, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(4, 400), nbins=st.sampled_from([2, 16, 257, 4096]))
def test_auroc_histogram_properties(seed, n, nbins):
    """0 <= AUROC <= 1; equals the oracle on the binned scores; complementing a two-class problem mirrors it around 1/2;
    a class whose scores all share one bin (or with no positives / negatives) gives 1/2 or 0."""
    from hss.sharding import auroc_from_histograms

    rng = np.random.default_rng(seed)
    target = rng.integers(0, 4, n)
    logits = rng.standard_normal((n, 4)).astype(np.float32) + 2.0 * np.eye(4, dtype=np.float32)[target] * rng.random((n, 1)).astype(np.float32)
    logp = torch.log_softmax(torch.from_numpy(logits), dim=1).numpy()
    h = torch.from_numpy(mo.histograms(logp, target, nbins))
    got = auroc_from_histograms(h)["auroc_per_class"].numpy()
    assert np.all((got >= 0) & (got <= 1))
    assert np.allclose(got, mo.auroc_binned(logp, target, nbins), atol=1e-12)
    # swap the roles of positives and negatives of every class: AUROC -> 1 - AUROC where it is defined
    swapped = auroc_from_histograms(h.flip(1))["auroc_per_class"].numpy()
    pos, neg = h[:, 1].sum(1).numpy(), h[:, 0].sum(1).numpy()
    ok = (pos > 0) & (neg > 0)
    assert np.allclose(swapped[ok], 1.0 - got[ok], atol=1e-12)
    assert np.all(got[~ok] == 0.0)


@settings(max_examples=40, deadline=None)
@given(B=st.integers(1, 6), T=st.integers(1, 9), H=st.integers(1, 5), seed=st.integers(0, 1000))
def test_row_shifted_recurrent_gradient_matches_the_concatenated_form(B, T, H, seed):
    """dG^T h_prev on row-shifted views + the B-row correction (hss/model/_train.py) == the product with h_prev built
    explicitly, for any batch / length / width and both directions (what autograd computes for nn.LSTM's weight_hh)."""
    from hss.model._train import edge_fixup, shifted_rows

    g = torch.Generator().manual_seed(seed)
    M = B * T
    dG = torch.randn(2, M, 4 * H, generator=g, dtype=torch.float64)
    out = torch.randn(B, T, 2 * H, generator=g, dtype=torch.float64)
    h0 = torch.randn(2, B, H, generator=g, dtype=torch.float64)
    o2 = out.reshape(M, 2 * H)
    hp = (torch.cat([h0[0].unsqueeze(1), out[:, :-1, :H]], dim=1).reshape(M, H),
          torch.cat([out[:, 1:, H:], h0[1].unsqueeze(1)], dim=1).reshape(M, H))
    for d in range(2):
        rg, ro, co = shifted_rows(M, H, d)
        edge_rows = torch.arange(B) * T + (0 if d == 0 else T - 1)
        got = dG[d][rg].t() @ o2[ro, co] + edge_fixup(dG[d][edge_rows], o2, h0[d], B, T, H, d)
        assert torch.allclose(got, dG[d].t() @ hp[d], rtol=1e-11, atol=1e-11)
