"""Parity of the FSST CUDA kernels (through the C-ABI) against the float64 oracle and the golden
vectors.  Tolerances (fp32 kernels vs float64 restatement, SURVEY 7 hard-part 2):
  * destination rows: fraction of (row, t) cells off by more than 1e-5*max|s| (= a value landed in
    a different row) must be < 2e-5;
  * agreeing cells: |delta| <= 2e-6 * max|s|;
  * stacked features: |delta| <= 2e-4 (z-scored units), abs magnitudes <= 2e-6 * max.
"""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import fsst_oracle as fo

pytestmark = pytest.mark.gpu

FS = 1000.0
W = fo.reference_window()
BAND = (25, 200)


@pytest.fixture(scope="module")
def lib(built):
    from hss import _lib

    assert torch.cuda.is_available()
    return _lib.lib()


def rows_and_values_ok(got, ref, frac=2e-5, tol=2e-6):
    scale = np.abs(ref).max()
    d = np.abs(got - ref)
    moved = d > 1e-5 * scale
    assert moved.mean() <= frac, f"{moved.sum()} of {moved.size} cells landed in another row"
    assert d[~moved].max() <= tol * scale


def test_three_kernels_through_the_c_abi(lib):
    """K1, K2, K3 called one by one on device pointers, each checked against the oracle."""
    from hss import _lib

    B, N = 3, 2000
    x = fo.synth_pcg_batch(B, N, seed=100)
    dg = fo.dtwin(W, FS)
    xd = torch.from_numpy(x).cuda()
    gd = torch.from_numpy(W.astype(np.float32)).cuda()
    dgd = torch.from_numpy(dg.astype(np.float32)).cuda()
    sg = torch.empty(B, 65, N, dtype=torch.complex64, device="cuda")
    sdg = torch.empty_like(sg)
    st = _lib.stream_ptr()
    _lib.check(lib.hssb_fsst_stft(xd.data_ptr(), B, N, gd.data_ptr(), dgd.data_ptr(), 128, sg.data_ptr(), sdg.data_ptr(), st), "stft")
    for b in range(B):
        rg, rdg = fo.stft_pair(x[b], FS, W)
        assert np.abs(sg[b].cpu().numpy() - rg[:65]).max() < 1e-6 * np.abs(rg).max()
        assert np.abs(sdg[b].cpu().numpy() - rdg[:65]).max() < 2e-6 * np.abs(rdg).max()
    # K2 full spectrum and band-fused
    t_full = torch.empty(B, 65, N, dtype=torch.complex64, device="cuda")
    _lib.check(lib.hssb_fsst_reassign(sg.data_ptr(), sdg.data_ptr(), B, N, 128, FS, 0, 64, t_full.data_ptr(), None, st), "reassign")
    t_band = torch.empty(B, 22, N, dtype=torch.complex64, device="cuda")
    stats = torch.zeros(lib.hssb_fsst_stats_words(B, N), dtype=torch.float64, device="cuda")
    _lib.check(lib.hssb_fsst_reassign(sg.data_ptr(), sdg.data_ptr(), B, N, 128, FS, 4, 25, t_band.data_ptr(), stats.data_ptr(), st), "reassign")
    assert torch.equal(t_band, t_full[:, 4:26])
    for b in range(B):
        s, _, _ = fo.fsst(x[b], FS, W)
        rows_and_values_ok(t_full[b].cpu().numpy(), s)
    # moment partials: merged by hand they must give the window's mean / M2
    parts = stats[: B * 16 * 6].cpu().numpy().reshape(B, 16, 2, 3)
    tb = t_band.cpu().numpy()
    for b in range(B):
        for c, v in enumerate((tb[b].real.astype(np.float64), tb[b].imag.astype(np.float64))):
            n = parts[b, :, c, 0].sum()
            mean = (parts[b, :, c, 0] * parts[b, :, c, 1]).sum() / n
            m2 = (parts[b, :, c, 2] + parts[b, :, c, 0] * (parts[b, :, c, 1] - mean) ** 2).sum()
            assert n == v.size and abs(mean - v.mean()) < 1e-9 and abs(m2 / (n - 1) - v.var(ddof=1)) < 1e-8 * v.var()
    # K3
    out = torch.empty(B, N, 44, dtype=torch.float32, device="cuda")
    _lib.check(lib.hssb_fsst_finish(t_band.data_ptr(), stats.data_ptr(), B, N, 22, 2, out.data_ptr(), st), "finish")
    for b in range(B):
        ref = fo.fsst_features(x[b], FS, W, stack=True, truncate_freq=BAND)
        assert np.abs(out[b].cpu().numpy() - ref).max() < 2e-4
    mag = torch.empty(B, N, 22, dtype=torch.float32, device="cuda")
    _lib.check(lib.hssb_fsst_finish(t_band.data_ptr(), None, B, N, 22, 1, mag.data_ptr(), st), "finish abs")
    assert torch.allclose(mag, t_band.abs().transpose(1, 2), rtol=1e-6, atol=0)


@pytest.mark.parametrize("name", ["fsst_pcg.npz", "fsst_noise.npz"])
def test_golden_vectors(lib, golden_dir, name):
    from hss.transforms import FSST

    g = np.load(os.path.join(golden_dir, name))
    x = torch.from_numpy(g["x"])
    trunc = tuple(g["truncate"])
    feats = FSST(float(g["fs"]), window=g["window"], truncate_freq=trunc, stack=True)(x)
    assert feats.shape == g["features"].shape and feats.dtype == torch.float32 and not feats.is_cuda
    assert np.abs(feats.numpy() - g["features"]).max() < 2e-4
    raw = FSST(float(g["fs"]), window=g["window"], truncate_freq=trunc)(x)
    assert raw.dtype == torch.complex64 and tuple(raw.shape) == g["s_band"].shape
    rows_and_values_ok(raw.numpy(), g["s_band"])
    mags = FSST(float(g["fs"]), window=g["window"], truncate_freq=trunc, abs=True)(x)
    assert np.abs(mags.numpy() - g["magnitudes"]).max() < 2e-6 * g["magnitudes"].max() + 1e-5 * (np.abs(raw.numpy() - g["s_band"]).max() > 0)


def test_wrapper_shapes_branches_and_input_forms(lib):
    """Branch precedence truncate -> abs -> stack -> raw, [N] / [N,1] / float64 inputs, CUDA in/out."""
    from hss.transforms import FSST

    x = torch.from_numpy(fo.synth_pcg(2000, seed=5))
    a = FSST(1000, window=W, truncate_freq=BAND, stack=True)(x)
    b = FSST(1000, window=W, truncate_freq=BAND, stack=True)(x.unsqueeze(1))          # dataset path: [2000, 1]
    c = FSST(1000, window=W, truncate_freq=BAND, stack=True)(x.double())              # visualize_signals.py:10
    assert a.shape == (2000, 44) and torch.equal(a, b) and torch.equal(a, c)
    both = FSST(1000, window=W, truncate_freq=BAND, abs=True, stack=True)(x)           # abs wins over stack
    assert both.shape == (2000, 22) and (both >= 0).all()
    full = FSST(1000, window=W)(x)
    assert full.shape == (65, 2000) and full.dtype == torch.complex64
    d = FSST(1000, window=W, truncate_freq=BAND, stack=True)(x.cuda())
    assert d.is_cuda and torch.equal(d.cpu(), a)
    batch = FSST(1000, window=W, truncate_freq=BAND, stack=True).batch(torch.stack([x, x * 2.0]))
    assert batch.shape == (2, 2000, 44) and torch.equal(batch[0], a)
    assert (batch[0] - batch[1]).abs().max() < 5e-5                                    # scale invariance


@pytest.mark.parametrize("n", [1, 31, 33, 127, 129, 640, 2001])
def test_ragged_lengths(lib, n):
    from hss.transforms import FSST

    x = fo.synth_pcg(n, seed=n)
    got = FSST(1000, window=W)(torch.from_numpy(x)).numpy()
    ref, _, _ = fo.fsst(x, FS, W)
    rows_and_values_ok(got, ref, frac=1e-3)
    if n > 1:
        feats = FSST(1000, window=W, truncate_freq=BAND, stack=True)(torch.from_numpy(x)).numpy()
        assert np.abs(feats - fo.fsst_features(x, FS, W, stack=True, truncate_freq=BAND)).max() < 5e-4


def test_empty_and_zero_inputs(lib):
    from hss.transforms import FSST

    f = FSST(1000, window=W, truncate_freq=BAND, stack=True)
    assert f.batch(torch.zeros(0, 2000)).shape == (0, 2000, 44)
    z = FSST(1000, window=W)(torch.zeros(300))
    assert not torch.isnan(z.real).any() and z.abs().max() == 0
    assert torch.isnan(f(torch.zeros(300))).all()          # 0/0 z-score, like the reference


def test_kaiser256_long_range_reassignment(lib):
    from hss.transforms import FSST

    w = np.kaiser(256, 10.0)
    n = 3000
    t = np.arange(n) / FS
    x = np.cos(2 * np.pi * (50 * t + 0.5 * 200 / t[-1] * t ** 2)).astype(np.float32)
    got = FSST(1000, window=w)(torch.from_numpy(x)).numpy()
    ref, f, _ = fo.fsst(x, FS, w)
    assert got.shape == (129, n)
    # with a long tapered window many bins are ~0 and their IF estimate is noise: compare where the
    # energy is (ridge) and bound the mass that landed elsewhere
    ridge = np.abs(ref).argmax(0)
    assert (np.abs(got).argmax(0)[200:-200] == ridge[200:-200]).mean() > 0.999
    assert np.abs(np.abs(got).sum() - np.abs(ref).sum()) < 1e-3 * np.abs(ref).sum()


def test_config2_size_properties(lib):
    """BASELINE config 2 (1024 x 2000): size-independent properties at full size + spot parity."""
    from hss.transforms import FSST

    B, N = 1024, 2000
    base = fo.synth_pcg_batch(16, N, seed=68)
    x = torch.from_numpy(np.tile(base, (B // 16, 1)) * np.linspace(0.5, 2.0, B, dtype=np.float32)[:, None]).cuda()
    raw = FSST(1000, window=W).batch(x)                                                # [B, 65, N]
    rec = 2 * raw.real.sum(1) - raw.real[:, 0] - raw.real[:, 64]                        # reconstruction identity
    assert (rec - 128 * float(W[64]) * x).abs().max() < 2e-3 * x.abs().max() * 128
    feats = FSST(1000, window=W, truncate_freq=BAND, stack=True).batch(x)
    assert feats.shape == (B, N, 44)
    for half in (feats[:, :, :22], feats[:, :, 22:]):
        assert half.mean(dim=(1, 2)).abs().max() < 1e-4
        assert (half.std(dim=(1, 2), unbiased=True) - 1).abs().max() < 1e-4
    # scale invariance across the tiled copies (same signal, different gain)
    assert (feats[0] - feats[16]).abs().max() < 2e-4
    for b in (0, 517, 1023):
        ref = fo.fsst_features(x[b].cpu().numpy(), FS, W, stack=True, truncate_freq=BAND)
        assert np.abs(feats[b].cpu().numpy() - ref).max() < 2e-4


def test_host_entry_point_matches_device_path(lib):
    """hssb_fsst_host: the call a binding replacing ssq.fsst would make (host buffers in/out)."""
    from hss import _lib
    from hss.transforms import FSST

    B, N = 2, 1500
    x = fo.synth_pcg_batch(B, N, seed=40)
    dg = fo.dtwin(W, FS)
    out = np.zeros((B, 65, N), dtype=np.complex64)
    rc = lib.hssb_fsst_host(x.ctypes.data, B, N, FS, W.ctypes.data, dg.ctypes.data, 128, 0, 64, 0, out.ctypes.data)
    _lib.check(rc, "hssb_fsst_host")
    dev = FSST(1000, window=W).batch(torch.from_numpy(x))
    assert np.array_equal(out, dev.numpy())
    feats = np.zeros((B, N, 44), dtype=np.float32)
    _lib.check(lib.hssb_fsst_host(x.ctypes.data, B, N, FS, W.ctypes.data, dg.ctypes.data, 128, 4, 25, 2, feats.ctypes.data), "host")
    assert np.array_equal(feats, FSST(1000, window=W, truncate_freq=BAND, stack=True).batch(torch.from_numpy(x)).numpy())


def test_frames_of_a_recording_equal_per_frame_calls(lib):
    """FSST.frames: one call over the strided frames of a recording == the per-frame loop of the reference dataset."""
    from hss.transforms import FSST
    from hss.utils.preprocess import frame_signal

    x = torch.from_numpy(fo.synth_pcg_batch(1, 9000, seed=5)[0])
    f = FSST(1000, window=fo.reference_window(), truncate_freq=(25, 200), stack=True)
    out = f.frames(x.cuda(), 1000, 2000)
    frames, _ = frame_signal(x, torch.zeros(x.shape[0], dtype=torch.int64), 1000, 2000)
    assert out.shape == (len(frames), 2000, 44) and len(frames) == 7
    w = fo.reference_window()
    for i, fr in enumerate(frames):
        # against the ORACLE on every frame (each frame zero-padded and z-scored on its own, heart_sounds.py:160-169) ...
        ref = fo.fsst_features(fr[:, 0].numpy(), 1000, w, stack=True, truncate_freq=(25, 200))
        assert np.abs(out[i].cpu().numpy() - ref).max() < 2e-4
        assert torch.equal(out[i].cpu(), f(fr))          # ... and the CPU-in -> CPU-out per-item path gives the same bits


def test_config5_long_windows_at_2khz(lib):
    """BASELINE config 5 geometry at reduced size: 2 kHz windows, band (50, 400) Hz -> rows 4..25 -> 44 features, long N."""
    from hss.transforms import FSST

    fs, N, B = 2000.0, 30000, 3
    x = torch.from_numpy(fo.synth_pcg_batch(B, N, fs=fs, seed=11))
    w = fo.reference_window()
    f = FSST(fs, window=w, truncate_freq=(50, 400), stack=True)
    out = f.batch(x.cuda()).cpu().numpy()
    assert out.shape == (B, N, 44)
    for b in range(B):
        ref = fo.fsst_features(x[b].numpy(), fs, w, stack=True, truncate_freq=(50, 400))
        assert np.abs(out[b] - ref).max() < 5e-4


def test_recording_to_frames_equals_the_dataset_loop(lib):
    """hss.utils.ingest.recording_to_frames == frame_signal(x, y - 1) + per-frame FSST of reference heart_sounds.py:155-169."""
    from hss.transforms import FSST
    from hss.utils import frame_signal, recording_to_frames
    from workloads import synthetic_targets

    x = torch.from_numpy(fo.synth_pcg_batch(1, 6500, seed=9)[0])
    y = torch.from_numpy(synthetic_targets(1, 6500)[0]) + 1
    f = FSST(1000, window=fo.reference_window(), truncate_freq=(25, 200), stack=True)
    feats, labels = recording_to_frames(x, y, f)
    frames, lab = frame_signal(x, y - 1, 1000, 2000)
    assert feats.shape == (len(frames), 2000, 44) and labels.shape == (len(frames), 2000) and feats.is_cuda
    w = fo.reference_window()
    for i, (fr, lb) in enumerate(zip(frames, lab)):
        ref = fo.fsst_features(fr[:, 0].numpy(), 1000, w, stack=True, truncate_freq=(25, 200))      # the oracle, per frame
        assert np.abs(feats[i].cpu().numpy() - ref).max() < 2e-4
        assert torch.equal(feats[i].cpu(), f(fr)) and torch.equal(labels[i].cpu(), lb[:, 0])
    short = recording_to_frames(x[:1500], y[:1500], f)
    assert short[0].shape[0] == 0 and short[1].shape[0] == 0


def test_config1_single_csv_recording_as_written(lib):
    """BASELINE config 1 as SURVEY 8d writes it: tests/data/0001.csv (35 000 rows, seed 68) -> 33 frames of 2000 at stride 1000
    (reference test/test_dataset.py:37,67-69) through (a) load_recording_csv -> recording_to_frames, one device call, and
    (b) the reference dataset's own loop -- _load_file, frame_signal(x, y - 1), per-frame transform on CPU tensors
    (heart_sounds.py:155-169,193-201).  Every frame is checked against the float64 oracle; (a) and (b) give the same bits."""
    import pandas as pd

    from hss.transforms import FSST
    from hss.utils import frame_signal, load_recording_csv, recording_to_frames

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "0001.csv")
    x, y = load_recording_csv(path)
    df = pd.read_csv(path, skiprows=1, names=["Signals", "Labels"])                    # heart_sounds.py:193-197
    assert torch.equal(x, torch.tensor(df.loc[:, "Signals"].to_numpy(), dtype=torch.float32))
    assert torch.equal(y, torch.tensor(df.loc[:, "Labels"].to_numpy(), dtype=torch.int64))
    assert x.shape == (35_000,) and int(y.min()) == 1 and int(y.max()) == 4

    w = fo.reference_window()
    f = FSST(1000, window=w, truncate_freq=(25, 200), stack=True)
    feats, labels = recording_to_frames(x, y, f)
    assert feats.shape == (33, 2000, 44) and labels.shape == (33, 2000) and labels.dtype == torch.int64
    frames, labs = frame_signal(x, y - 1, 1000, 2000)
    assert len(frames) == 33
    worst = 0.0
    for i, (fr, lb) in enumerate(zip(frames, labs)):
        item = f(fr)                                                                    # CPU tensor [2000, 1] in, CPU [2000, 44] out
        assert item.shape == (2000, 44) and not item.is_cuda and lb.squeeze(1).shape == (2000,)
        ref = fo.fsst_features(fr[:, 0].numpy(), 1000, w, stack=True, truncate_freq=(25, 200))
        worst = max(worst, float(np.abs(item.numpy() - ref).max()))
        assert torch.equal(feats[i].cpu(), item) and torch.equal(labels[i].cpu(), lb.squeeze(1))
    assert worst < 2e-4, worst
    print("config 1: 33 frames, max |feature - oracle| =", worst)


def test_second_device_after_first(lib):
    """Function attributes (opt-in shared memory) and cluster occupancy are per device: FSST (nwin 128 and 256) and the model
    on cuda:1 after cuda:0 in one process give the same results."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from hss.model.segmenter import HeartSoundSegmenter
    from hss.transforms import FSST

    x = torch.from_numpy(fo.synth_pcg_batch(3, 2000, seed=3))
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        with torch.cuda.device(dev):
            f128 = FSST(1000, window=fo.reference_window(), truncate_freq=(25, 200), stack=True)
            f256 = FSST(1000, window=np.kaiser(256, 10.0))
            torch.manual_seed(1)
            m = HeartSoundSegmenter(input_size=44, batch_size=3).eval()
            feats = f128.batch(x.to(dev))
            outs.append((feats.cpu(), f256.batch(x.to(dev)).cpu(), m(feats).cpu()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_stream_recordings_pipeline_equals_per_recording_path(lib, tmp_path):
    """hss.utils.ingest.stream_recordings: native multi-file parse -> pinned staging -> async H2D on a copy stream
    -> FSST, pipelined over groups of files, gives the same tensors as load_recording_csv + recording_to_frames per file;
    recordings shorter than one frame are skipped (heart_sounds.py:160-161)."""
    from hss.transforms import FSST
    from hss.utils import load_recording_csv, recording_to_frames, stream_recordings
    from workloads import synth_pcg, synthetic_targets

    lens = [4100, 2000, 1500, 9000, 3050, 2500, 7777]
    paths = []
    for i, n in enumerate(lens):
        x = synth_pcg(n, 1000.0, 100 + i)
        y = synthetic_targets(1, n)[0] + 1
        path = tmp_path / f"{i:04d}.csv"
        with open(path, "w") as f:
            f.write("Signals,Labels\n")
            for a, b in zip(x, y):
                f.write(f"{float(a):.9g},{int(b)}\n")
        paths.append(str(path))
    f = FSST(1000, window=fo.reference_window(), truncate_freq=(25, 200), stack=True)
    streamed = list(stream_recordings(paths, f, group=3, threads=2))
    expected = []
    for p in paths:
        x, y = load_recording_csv(p)
        feats, labels = recording_to_frames(x, y, f)
        if feats.shape[0]:
            expected.append((feats, labels))
    assert len(streamed) == len(expected) == 6                    # only the 1500-sample recording is shorter than one frame
    for (a, la), (b, lb) in zip(streamed, expected):
        assert a.shape == b.shape and torch.equal(a, b) and torch.equal(la, lb)


@pytest.mark.parametrize("nwin,beta", [(64, 4.0), (100, 6.0), (101, 6.0), (255, 10.0), (512, 10.0)])
def test_generic_window_lengths(lib, nwin, beta):
    """ssq.fsst takes any window (reference synchrosqueeze.py:48): lengths other than 128 / 256 -- even and odd -- run on the
    direct-DFT STFT kernel and the modulo-nfft reassignment kernel; same checks against the float64 oracle as the radix kernels."""
    from hss.transforms import FSST

    fs, N, B = 1000.0, 1500, 3
    w = np.kaiser(nwin, beta)
    x = fo.synth_pcg_batch(B, N, seed=nwin)
    raw = FSST(fs, window=w).batch(torch.from_numpy(x).cuda()).cpu().numpy()
    assert raw.shape == (B, nwin // 2 + 1, N)
    for b in range(B):
        s, _, _ = fo.fsst(x[b], fs, w)
        rows_and_values_ok(raw[b], s, frac=1e-4, tol=4e-6)
    band = (40, 180)
    feats = FSST(fs, window=w, truncate_freq=band, stack=True).batch(torch.from_numpy(x).cuda()).cpu().numpy()
    mags = FSST(fs, window=w, truncate_freq=band, abs=True)(torch.from_numpy(x[0]))
    for b in range(B):
        ref = fo.fsst_features(x[b], fs, w, stack=True, truncate_freq=band)
        assert feats[b].shape == ref.shape and np.abs(feats[b] - ref).max() < 5e-4
    ref_m = fo.fsst_features(x[0], fs, w, abs=True, truncate_freq=band)
    assert mags.shape == ref_m.shape and (np.abs(mags.numpy() - ref_m) > 1e-5 * ref_m.max()).mean() < 1e-4


def test_closed_form_properties_on_the_cuda_output(lib):
    """Independent of the oracle (which restates MATLAB fsst and cannot be pinned against libssq): closed-form facts asserted
    directly on the CUDA transform, for the reference's Kaiser(128, 0.5) window and MATLAB's default Kaiser(256, 10).

    1. Reconstruction identity.  Reassignment only moves S_g[k, t] e^{-i pi k} between rows of a column, so the two-sided column
       sum is the inverse DFT of the windowed frame at its centre sample: sum_k T[k, t] = nfft * g[nwin/2] * x[t] for EVERY t,
       i.e. with the one-sided rows  2 Re sum_{k=0}^{K-1} T - Re T[0] - Re T[K-1] = nfft * g[nwin/2] * x[t].
    2. Linear chirp cos(2 pi (f0 t + c t^2 / 2)): the ridge of column t sits at row round((f0 + c t) nfft / fs), and holds
       most of the column's energy (steps 4-6: IF estimate, phase shift, frequency reassignment).
    3. Two bin-centred tones of amplitudes 1 and 0.5: interior columns have |T[k1]| = nfft g[nwin/2] / 2 and half of that at k2."""
    from hss.transforms import FSST

    fs, N = 1000.0, 2000
    t = np.arange(N) / fs
    for nwin, beta in ((128, 0.5), (256, 10.0)):
        w = np.kaiser(nwin, beta)
        K, g_c = nwin // 2 + 1, w[nwin // 2]
        f = FSST(fs, window=w)
        # 1. identity on a PCG-like signal and on noise
        x = np.stack([fo.synth_pcg(N, seed=3), np.random.default_rng(1).standard_normal(N).astype(np.float32)])
        T = f.batch(torch.from_numpy(x).cuda()).cpu().numpy().astype(np.complex128)
        lhs = 2 * T.real.sum(axis=1) - T[:, 0].real - T[:, K - 1].real
        assert np.abs(lhs - nwin * g_c * x).max() < 2e-5 * nwin * np.abs(x).max()
        # 2. chirp 100 -> 300 Hz
        f0, c = 100.0, 100.0
        chirp = np.cos(2 * np.pi * (f0 * t + 0.5 * c * t * t)).astype(np.float32)
        Tc = np.abs(f(torch.from_numpy(chirp)).numpy())
        inner = slice(nwin, N - nwin)
        ridge = Tc[:, inner].argmax(axis=0)
        expect = np.rint((f0 + c * t[inner]) * nwin / fs).astype(int)
        assert np.abs(ridge - expect).max() <= 1 and (ridge == expect).mean() > 0.9
        if beta >= 4.0:          # (a well-concentrated window; the reference's near-rectangular Kaiser(0.5) leaks by design)
            e = Tc[:, inner] ** 2
            near = sum(np.take_along_axis(e, np.clip(expect + d, 0, K - 1)[None, :], axis=0)[0] for d in (-1, 0, 1))
            assert (near / e.sum(axis=0)).min() > 0.9
        # 3. two bin-centred tones
        k1, k2 = nwin // 8, nwin // 4 + 3
        tone = (np.cos(2 * np.pi * k1 * fs / nwin * t) + 0.5 * np.cos(2 * np.pi * k2 * fs / nwin * t)).astype(np.float32)
        Tt = np.abs(f(torch.from_numpy(tone)).numpy())[:, inner]
        assert np.abs(Tt[k1] - nwin * g_c / 2).max() < 5e-3 * nwin * g_c / 2
        assert np.abs(Tt[k2] - nwin * g_c / 4).max() < 1e-2 * nwin * g_c / 4
        rest = np.delete(Tt, [k1, k2], axis=0)
        assert rest.max() < 2e-2 * nwin * g_c / 2
