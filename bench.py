#!/usr/bin/env python
"""bench.py -- PCG samples/s through FSST + BiLSTM on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                          # the reference algorithm on host cores

One step = one pass of the hot path over one batch of synthetic PCG windows per GPU:
FSST -> BiLSTM segmenter (eval forward) -> log-probabilities + argmax labels -> metric state (4x4 confusion counts,
loss sum, count: 18 scalars accumulated on the device).  The ONLY collective is one NCCL all-reduce of those 18
scalars after the last step (north star: "final metric all-reduce").  Weak scaling: every rank owns WINDOWS_PER_GPU
windows (BASELINE config 4's shard: 4096 windows / 8 GPUs = 512 per GPU), taken from ONE seeded set of 4096 windows
and ONE [2, 4096, 240] draw of h0 / c0, so 1-GPU and N-GPU runs see the same data (`config4_confusion`).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import os
import sys

if ("--impl" in sys.argv and "reference" in sys.argv) or "--impl=reference" in sys.argv:
    # the reference arm uses every host core: torchrun exports OMP_NUM_THREADS=1 to its children, and the OpenMP / MKL
    # runtimes read the environment when they are loaded -- so this happens before numpy / torch are imported
    for _k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[_k] = str(os.cpu_count() or 1)

import argparse  # noqa: E402
import gc  # noqa: E402
import json  # noqa: E402
import subprocess  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402
import zlib  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "heart-sounds-segmentation_b200"))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

FS = 1000.0
N_SAMPLES = 2000            # 2 s windows at 1 kHz (reference hss/datasets/heart_sounds.py:123-124)
NWIN = 128
BAND = (25, 200)
K_BINS, KT = 65, 22
WINDOWS_PER_GPU = 512       # BASELINE config 4 shard
CONFIG4_WINDOWS = 4096      # BASELINE config 4: the global batch
# useful FLOPs (1x, the fp32 contraction) per PCG sample of every launch of the kernel in one step
FLOP_PER_SAMPLE = {"tc_inproj_l0": 2 * 44 * 1920.0, "tc_inproj_l1": 2 * 480 * 1920.0, "simt_inproj": 2 * (44 + 480) * 1920.0,
                   "tc_recurrent": 2 * 2 * 240 * 1920.0, "tc_recurrent_l1": 2 * 240 * 1920.0, "tc_recurrent_l2": 2 * 240 * 1920.0,
                   "simt_recurrent": 2 * 2 * 240 * 1920.0}
# algorithmic HBM bytes per PCG sample: K1 4 + 2*65*8, K2 2*65*8 + 22*8, K3 352; the fused STFT+reassign kernel
# reads x and writes the band rows only
BYTES_PER_SAMPLE = {"stft_hop1": 4 + 2 * K_BINS * 8, "if_reassign": 2 * K_BINS * 8 + KT * 8, "normalise": 16 * KT,
                    "fsst_fused": 4 + KT * 8, "head": 2048 + 20}
FSST_KERNELS = ("stft_hop1", "if_reassign", "fsst_fused", "stats_finalize", "normalise")
FSST_BOUND_BYTES = 4 + 8 * KT   # what end-to-end "FSST GB/s" is quoted against (x in, 44 features out)


def workload_name(windows: int) -> str:
    return (f"config 4 shard: FSST(kaiser128,25-200Hz,stack)+BiLSTM(44->240x2x2->4), {windows} windows x {N_SAMPLES} "
            "samples per GPU")


def bench_config(windows: int) -> dict:
    """The `config` object: identical for this repo's arm and for the reference arm."""
    return {"workload": workload_name(windows), "windows_per_gpu": windows, "samples_per_window": N_SAMPLES}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tensor_tflops": p["bf16_tflops_sustained"], "tensor_tflops_burst": p["bf16_tflops"],
                "source": "MEASURED_PEAKS.json (of measured)"}
    return {"hbm_gbs": 6650.0, "tensor_tflops": 1400.0, "tensor_tflops_burst": 1590.0, "source": "B200_PROFILING.md fallback (of fallback)"}


TRAFFIC_FILES = ("r02_traffic.json", "r01d_traffic.json")


def load_traffic():
    for name in TRAFFIC_FILES:
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            try:
                return json.load(open(path))["kernels"], os.path.join("profiles", name)
            except (ValueError, KeyError):
                pass
    return {}, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc, self.mark_at = index, [], None, 0

    def mark(self):
        """Samples from here on belong to the timed region (the process is started before the warm-up: nvidia-smi can take
        longer to come up than a short timed region lasts)."""
        self.mark_at = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        lines = self.lines[self.mark_at:] if len(self.lines) > self.mark_at else self.lines   # else: warm-up samples, also under load
        for ln in lines:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# synthetic workload (same generators as the parity tests)
# --------------------------------------------------------------------------------------------------
def config4_windows() -> np.ndarray:
    """The 4096 windows of BASELINE config 4 (one seeded set; ranks take contiguous blocks of it)."""
    from workloads import tiled_windows

    return tiled_windows(CONFIG4_WINDOWS, N_SAMPLES, FS, 68)


def shard_windows(all_windows: np.ndarray, shard: int, per: int) -> np.ndarray:
    idx = (np.arange(per) + shard * per) % all_windows.shape[0]
    return np.ascontiguousarray(all_windows[idx])


def synthetic_targets(n_windows: int, n: int = N_SAMPLES) -> np.ndarray:
    from workloads import synthetic_targets as st

    return st(n_windows, n)


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: restated FSST (C, OpenMP) + torch-CPU BiLSTM on the host cores
# --------------------------------------------------------------------------------------------------
class CpuReference:
    def __init__(self, batch: int):
        import ctypes
        from oracle import fsst_oracle as fo
        from oracle import lstm_oracle as lo

        self.fo, self.lo, self.batch = fo, lo, batch
        so = os.path.join(ROOT, "oracle", "_build", "libhss_oracle.so")
        if not os.path.exists(so):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
        self.lib = ctypes.CDLL(so)
        P = ctypes.c_void_p
        self.lib.hsso_fsst_features.argtypes = [P, ctypes.c_int64, ctypes.c_int64, ctypes.c_double, P, P, ctypes.c_int,
                                                ctypes.c_int, ctypes.c_int, ctypes.c_int, P]
        # every host core, set explicitly (torchrun exports OMP_NUM_THREADS=1)
        cores = os.cpu_count() or 1
        if hasattr(self.lib, "hsso_set_num_threads"):
            self.lib.hsso_set_num_threads.argtypes = [ctypes.c_int]
            self.lib.hsso_set_num_threads(cores)
        torch.set_num_threads(cores)
        self.window = fo.reference_window(NWIN)
        self.dwindow = fo.dtwin(self.window, FS)
        self.fsst_threads = int(self.lib.hsso_num_threads())
        self.torch_threads = torch.get_num_threads()
        self.params = None

    def prepare(self, batch: int):
        self.params, self.h0, self.c0 = self.lo.reference_params(68, 2 * KT, batch, 240)

    def step(self, x: np.ndarray):
        if self.params is None or self.h0.shape[1] != x.shape[0]:
            self.prepare(x.shape[0])
        feats = np.empty((x.shape[0], N_SAMPLES, 2 * KT), dtype=np.float32)
        t0 = time.perf_counter()
        rc = self.lib.hsso_fsst_features(x.ctypes.data, x.shape[0], N_SAMPLES, FS, self.window.ctypes.data,
                                         self.dwindow.ctypes.data, NWIN, 4, 25, 2, feats.ctypes.data)
        assert rc == 0
        t1 = time.perf_counter()
        logp = self.lo.forward_torch(self.params, self.h0, self.c0, torch.from_numpy(feats))
        labels = logp.argmax(-1)
        t2 = time.perf_counter()
        return labels, t1 - t0, t2 - t1

    def describe(self, batch: int) -> str:
        return (f"{batch} windows x {N_SAMPLES} samples per step; FSST = C/OpenMP restatement of MATLAB fsst ({self.fsst_threads} "
                f"threads; libssq itself is unobtainable), BiLSTM = torch-CPU nn.LSTM restatement of segmenter.py "
                f"({self.torch_threads} threads)")


def run_reference(args, rank: int):
    if rank != 0:
        return
    # the same windows per step as this repo's arm (config 4's shard), unless --ref-batch bounds the sample further
    batch = args.ref_batch if args.ref_batch > 0 else args.windows
    ref = CpuReference(batch)
    x = shard_windows(config4_windows(), 0, batch)
    ref.step(x[:min(batch, 32)])          # first-touch / thread-pool warm-up on a small block
    for _ in range(args.warmup):
        ref.step(x)
    t_f = t_l = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, a, b = ref.step(x)
        t_f += a; t_l += b
    dt = time.perf_counter() - t0
    value = batch * N_SAMPLES * args.steps / dt
    line = {
        "impl": "reference", "metric": "PCG samples/s through FSST+BiLSTM", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args.windows),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port", "sample": ref.describe(batch),
                         "fsst_samples_per_s": batch * N_SAMPLES * args.steps / t_f, "lstm_samples_per_s": batch * N_SAMPLES * args.steps / t_l},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------------------
def timed(fn, steps: int, warmup: int, flush=None):
    """ms per call of fn() (CUDA events, L2 flushed between calls)."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(steps):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        total += a.elapsed_time(b)
    return total / steps


def side_configs(dev, flush, peaks, _lib) -> dict:
    """The other single-GPU configurations of BASELINE.json as sub-records of the one JSON line (N = 1 only)."""
    from hss.model.segmenter import HeartSoundSegmenter
    from hss.transforms import FSST
    from workloads import reference_window, synth_pcg_batch, tiled_windows

    out = {}
    fsst = FSST(FS, window=reference_window(NWIN), truncate_freq=BAND, stack=True)
    # ---- config 2: FSST only, 1024 windows ----
    B2 = 1024
    x2 = torch.from_numpy(tiled_windows(B2, N_SAMPLES, FS, 68)).to(dev)
    _lib.prof_enable(True)
    fsst.batch(x2)
    torch.cuda.synchronize()
    _lib.prof_read()
    steps = 5
    ms = timed(lambda: fsst.batch(x2), steps, 0, flush)
    prof = _lib.prof_read()
    _lib.prof_enable(False)
    units = B2 * N_SAMPLES
    kern = {}
    for name, (cnt, tot) in prof.items():
        e = {"ms": tot / cnt}
        if name in BYTES_PER_SAMPLE:
            gbs = BYTES_PER_SAMPLE[name] * units / (tot / cnt * 1e-3) / 1e9
            e.update({"GB/s": gbs, "frac_of_hbm_peak": gbs / peaks["hbm_gbs"]})
        kern[name] = e
    e2e_gbs = FSST_BOUND_BYTES * units / (ms * 1e-3) / 1e9
    out["config2_fsst_only"] = {
        "workload": f"FSST only, {B2} windows x {N_SAMPLES} samples, device resident", "ms": ms, "samples_per_s": units / (ms * 1e-3),
        "kernels": kern, "end_to_end_GB/s_vs_180B_bound": e2e_gbs, "end_to_end_frac_of_hbm_peak": e2e_gbs / peaks["hbm_gbs"],
        "note": "end-to-end figure = (4 B in + 176 B out per sample) / total FSST time: the fully fused lower bound of SURVEY 8d; the FSST "
                "is FFT-issue bound, not HBM bound, against that bound"}
    del x2
    # ---- config 3: FSST + BiLSTM, batch 50 ----
    B3 = 50
    x3 = torch.from_numpy(synth_pcg_batch(B3, N_SAMPLES, FS, 68)).to(dev)
    torch.manual_seed(68)
    m3 = HeartSoundSegmenter(input_size=2 * KT, batch_size=B3).eval()
    ms = timed(lambda: m3.forward_with_labels(fsst.batch(x3)), 5, 2, flush)
    out["config3_batch50"] = {"workload": f"FSST + BiLSTM, batch {B3} x {N_SAMPLES} samples", "ms": ms, "samples_per_s": B3 * N_SAMPLES / (ms * 1e-3)}
    del m3
    # ---- config 3, training: one optimisation step (forward in training mode, loss, backward, clip + Adam) at batch 50 ----
    from hss.optim import ClipAdam

    torch.manual_seed(68)
    mt = HeartSoundSegmenter(input_size=2 * KT, batch_size=B3).to(dev).train()
    opt = ClipAdam(mt.parameters(), lr=0.01, max_norm=1.0)
    feats = fsst.batch(x3)
    y3 = torch.from_numpy(synthetic_targets(B3)).to(dev)

    def train_step():
        opt.zero_grad(set_to_none=True)
        loss, _ = mt.training_loss(feats, y3)
        loss.backward()
        opt.step()

    ms = timed(train_step, 5, 3, flush)
    out["config3_training_step"] = {
        "workload": f"training step (main.py:67-82, 130-135): batch {B3} x {N_SAMPLES} samples, dropout 0.2, CE loss, clip 1.0 + Adam",
        "ms": ms, "samples_per_s": B3 * N_SAMPLES / (ms * 1e-3),
        "kernels": "projection + recurrences forward (K4, K5m TRAIN) and back-propagation through time (K5b) on tcgen05; weight / input "
                   "gradient GEMMs as 3 x TF32 library GEMMs on split operands; fused head + loss and clip + Adam kernels"}
    del mt, opt, feats, y3, x3
    # ---- config 5 shard: 64 windows x 120 000 samples @ 2 kHz ----
    B5, N5, fs5 = 64, 120_000, 2000.0
    f5 = FSST(fs5, window=reference_window(NWIN), truncate_freq=(50, 400), stack=True)
    base = synth_pcg_batch(4, N5, fs5, 5)
    x5 = torch.from_numpy(np.tile(base, (B5 // 4, 1)) * np.linspace(0.5, 2.0, B5, dtype=np.float32)[:, None]).to(dev)
    torch.manual_seed(5)
    m5 = HeartSoundSegmenter(input_size=2 * KT, batch_size=B5).eval()
    ms = timed(lambda: m5.forward_with_labels(f5.batch(x5)), 2, 1, flush)
    out["config5_shard"] = {"workload": f"FSST + BiLSTM, {B5} windows x {N5} samples @ 2 kHz (band 50-400 Hz -> 44 features), T = {N5} dependent steps",
                            "ms": ms, "samples_per_s": B5 * N5 / (ms * 1e-3)}
    del m5, x5, f5
    torch.cuda.empty_cache()
    return out


def run_ours(args, rank: int, local_rank: int, world: int):
    import torch.distributed as dist
    from hss import _lib
    from hss.model.segmenter import HeartSoundSegmenter
    from hss.sharding import allreduce_counts, metric_state, metrics_from_state
    from hss.transforms import FSST
    from workloads import reference_window

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.windows
    peaks = load_peaks()
    n_shards = max(1, CONFIG4_WINDOWS // B)

    all_windows = config4_windows()
    x_host = torch.from_numpy(shard_windows(all_windows, rank % n_shards, B)).pin_memory()
    y_dev = torch.from_numpy(synthetic_targets(B)).to(dev)
    x_dev = x_host.to(dev)
    fsst = FSST(FS, window=reference_window(NWIN), truncate_freq=BAND, stack=True)
    # weights as config 3 / 4 name them (seed 68 -> reference ctor); h0 / c0 of the whole job drawn once, sliced per shard
    torch.manual_seed(68)
    model = HeartSoundSegmenter(input_size=2 * KT, batch_size=B).eval()
    g = torch.Generator().manual_seed(4096)
    H0 = torch.randn(2, n_shards * B, 240, generator=g)
    C0 = torch.randn(2, n_shards * B, 240, generator=g)

    def use_shard(s):
        model.h0 = H0[:, s * B:(s + 1) * B].contiguous()
        model.c0 = C0[:, s * B:(s + 1) * B].contiguous()

    use_shard(rank % n_shards)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    from hss.pipeline import SegmentationPipeline

    pipe = SegmentationPipeline(fsst, model, dev)

    def step(x, state, x_next=None):
        """One pass of the hot path over one batch.  x_next (the following step's input, when the loop has it): its FSST is
        enqueued on a second stream and runs under this batch's recurrences (hss.pipeline) -- same kernels, same results."""
        if args.pipeline:
            logp, labels = pipe(x, prefetch=x_next)
        else:
            logp, labels = model.forward_with_labels(fsst.batch(x))
        metric_state(logp, y_dev, labels=labels, state=state)
        return labels

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    state = torch.zeros(18, dtype=torch.float64, device=dev)
    n_warm = max(args.warmup, 3)
    for i in range(n_warm):                   # the same call pattern as the timed steps (prefetch of the following batch)
        step(x_dev, state, x_dev if i + 1 < n_warm else None)
    pipe.close()
    allreduce_counts(state.clone())          # communicator warm-up
    barrier()

    # ---- device-resident timing: K steps, CUDA events per step, L2 flushed between steps; the metric all-reduce once at the end ----
    sampler.mark()
    _lib.prof_enable(True)
    _lib.prof_read()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps + 1)]
    state.zero_()
    gc.collect()        # (before the barrier: a gen-2 collection takes 10-50 ms and differs per rank)
    gc.disable()
    barrier()
    for i, (a, b) in enumerate(ev[:-1]):
        flush.zero_()
        a.record()
        step(x_dev, state, x_dev if i + 1 < args.steps else None)
        b.record()
    ev[-1][0].record()
    pipe.close()
    allreduce_counts(state)                   # the one collective of the job: 18 scalars
    ev[-1][1].record()
    barrier()
    gc.enable()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    ms_allreduce = ev[-1][0].elapsed_time(ev[-1][1])
    prof = _lib.prof_read()
    _lib.prof_enable(False)
    clocks = sampler.stop()
    final_metrics = metrics_from_state(state)

    # ---- the same K steps as the plain sequence model(fsst.batch(x)) (no overlap between batches), for comparison only ----
    ms_sequential, ms_sequential_steps = None, None
    if args.pipeline:
        scratch = torch.zeros(18, dtype=torch.float64, device=dev)
        seq_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        model.forward_with_labels(fsst.batch(x_dev))          # (one untimed pass: this call pattern allocates differently)
        gc.collect()
        gc.disable()                                           # as in the timed loop: a gen-2 collection stalls the enqueue for 10-100 ms
        barrier()
        for a, b in seq_ev:
            flush.zero_()
            a.record()
            lp, lb = model.forward_with_labels(fsst.batch(x_dev))
            metric_state(lp, y_dev, labels=lb, state=scratch)
            b.record()
        barrier()
        gc.enable()
        ms_sequential = sum(a.elapsed_time(b) for a, b in seq_ev) / args.steps
        ms_sequential_steps = [round(a.elapsed_time(b), 3) for a, b in seq_ev]

    # ---- end to end: pinned host input -> H2D -> path -> labels + metric state back on the host ----
    # Double buffered like a production ingest loop: step i is enqueued (H2D copy, kernels, D2H copies into pinned slot i % 2),
    # then the host waits for step i-1's event and reads its results -- every step's labels and counters reach the host inside
    # the timed region, one step behind the device.  Python GC off while timing (a gen-2 collection costs 30-90 ms).
    lab_slots = [torch.empty((B, N_SAMPLES), dtype=torch.int32).pin_memory() for _ in range(2)]
    st_slots = [torch.empty(18, dtype=torch.float64).pin_memory() for _ in range(2)]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    x_in = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
    state_e = torch.zeros(18, dtype=torch.float64, device=dev)

    def e2e_enqueue(i, n):
        # the H2D copy of step i + 1 is issued before step i's kernels so that its transform can be prefetched under them
        if i == 0:
            x_in[0].copy_(x_host, non_blocking=True)
        nxt = None
        if i + 1 < n:
            nxt = x_in[(i + 1) & 1]
            nxt.copy_(x_host, non_blocking=True)
        labels = step(x_in[i & 1], state_e, nxt)
        lab_slots[i & 1].copy_(labels, non_blocking=True)
        st_slots[i & 1].copy_(state_e, non_blocking=True)
        done[i & 1].record()

    def e2e_collect(i):
        done[i & 1].synchronize()
        return float(st_slots[i & 1][17]) + int(lab_slots[i & 1][0, 0])      # the host reads the step's results

    def e2e_run(n):
        seen = 0.0
        for i in range(n):
            e2e_enqueue(i, n)
            if i:
                seen += e2e_collect(i - 1)
        seen += e2e_collect(n - 1)
        pipe.close()
        allreduce_counts(state_e)
        return seen

    e2e_run(max(args.warmup, 3))
    state_e.zero_()
    gc.collect()
    gc.disable()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    gc.enable()
    ms_e2e = e0.elapsed_time(e1)

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    samples = B * N_SAMPLES * world * args.steps
    value = samples / (ms * 1e-3)

    # ---- 1-vs-N check (untimed): the metric state over the SAME 4096 windows of config 4, whatever N is ----
    # every rank takes shards rank, rank + world, ...; the all-reduced confusion counts must not depend on N
    cfg4 = torch.zeros(18, dtype=torch.float64, device=dev)
    if B * n_shards == CONFIG4_WINDOWS:
        for s in range(rank, n_shards, world):
            use_shard(s)
            xs = torch.from_numpy(shard_windows(all_windows, s, B)).to(dev)
            step(xs, cfg4)
        allreduce_counts(cfg4)
        use_shard(rank % n_shards)
    cfg4_host = cfg4.cpu()

    if rank == 0:
        per_kernel = {}
        launches = 0
        for name, (cnt, tot) in prof.items():
            launches += cnt
            per_launch_ms = tot / cnt
            units = B * N_SAMPLES          # samples processed by the launches of one step
            n_per_step = cnt / args.steps
            entry = {"launches_per_step": n_per_step, "ms_per_step": tot / args.steps}
            if name in BYTES_PER_SAMPLE:
                gbs = BYTES_PER_SAMPLE[name] * units / (per_launch_ms * 1e-3) / 1e9
                entry.update({"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"]})
            elif name == "tc_inproj_l1" and ("tc_inproj_l1_tail" in prof or "tc_inproj_l1_mid" in prof):
                entry["note"] = "the part of the layer-2 projection that runs alone on the GPU (outer time tiles); the rest is in _mid / _tail"
            elif name in FLOP_PER_SAMPLE:
                flop = FLOP_PER_SAMPLE[name]
                if name in ("tc_recurrent", "tc_recurrent_l1") and "tc_inproj_l0" not in prof:
                    flop += FLOP_PER_SAMPLE["tc_inproj_l0"]        # layer 1's input projection is fused into the recurrence kernel
                tf = flop * units / (tot / args.steps * 1e-3) / 1e12
                entry.update({"bound": "tensor", "achieved": tf, "peak": peaks["tensor_tflops"], "unit": "TFLOP/s", "frac": tf / peaks["tensor_tflops"],
                              "note": "useful FLOPs (1x) over all launches of this kernel in a step"})
            per_kernel[name] = entry
        # the two layers' recurrence launches are timed separately; `tc_recurrent` (both together) stays the roofline kernel
        if "tc_recurrent_l1" in per_kernel and "tc_recurrent_l2" in per_kernel:
            a, b = per_kernel["tc_recurrent_l1"], per_kernel["tc_recurrent_l2"]
            ms_both = a["ms_per_step"] + b["ms_per_step"]
            flop = FLOP_PER_SAMPLE["tc_recurrent"] + (FLOP_PER_SAMPLE["tc_inproj_l0"] if "tc_inproj_l0" not in prof else 0.0)
            tf = flop * B * N_SAMPLES / (ms_both * 1e-3) / 1e12
            per_kernel["tc_recurrent"] = {"launches_per_step": a["launches_per_step"] + b["launches_per_step"], "ms_per_step": ms_both, "bound": "tensor",
                                          "achieved": tf, "peak": peaks["tensor_tflops"], "unit": "TFLOP/s", "frac": tf / peaks["tensor_tflops"],
                                          "note": "both layers' recurrence launches together (layer 2 runs concurrently with the tail of the projection GEMM)"}
        # the layer-2 projection runs as up to three launches (tc_inproj_l1 alone on the GPU, _mid under the layer-1 recurrence,
        # _tail under the layer-2 recurrence): their times overlap with the recurrences', so the per-kernel times do not add up to the step
        # ncu evidence committed under profiles/: DRAM bytes per launch and tensor-pipe activity of every kernel of this workload;
        # only attached when the workload is the profiled one
        traffic, traffic_src = load_traffic() if (B == WINDOWS_PER_GPU) else ({}, None)
        if "tc_recurrent_l1" in traffic and "tc_recurrent_l2" in traffic and "tc_recurrent" not in traffic:
            a, b = traffic["tc_recurrent_l1"], traffic["tc_recurrent_l2"]
            ta, tb = a["ncu_ms_per_launch"], b["ncu_ms_per_launch"]
            traffic["tc_recurrent"] = {"dram_bytes_per_launch": a["dram_bytes_per_launch"] + b["dram_bytes_per_launch"],
                                       "tensor_pipe_active_pct": (a["tensor_pipe_active_pct"] * ta + b["tensor_pipe_active_pct"] * tb) / (ta + tb)}
        for name, entry in per_kernel.items():
            tr = traffic.get(name)
            if tr:
                entry["traffic"] = tr["dram_bytes_per_launch"]
                if tr["tensor_pipe_active_pct"] > 0:
                    entry["ncu_tensor_pipe_active_pct"] = tr["tensor_pipe_active_pct"]
        dominant = max((k for k in per_kernel if k not in ("tc_recurrent_l1", "tc_recurrent_l2")), key=lambda k: per_kernel[k]["ms_per_step"])
        d = per_kernel[dominant]
        roofline = {"kernel": dominant, "bound": d.get("bound"), "achieved": d.get("achieved"), "peak": d.get("peak"), "unit": d.get("unit"),
                    "frac": d.get("frac"), "traffic": d.get("traffic"), "share_of_step": d["ms_per_step"] / (ms / args.steps),
                    "peak_source": peaks["source"],
                    "note": "achieved = useful fp32-equivalent FLOPs (1x) of the W_hh contraction / launch time; the kernel issues 3 fp16 MMAs "
                            "per product on 8-CTA clusters, ncu sm__pipe_tensor_cycles_active in ncu_tensor_pipe_active_pct"
                            if dominant == "tc_recurrent" else None,
                    "ncu_tensor_pipe_active_pct": d.get("ncu_tensor_pipe_active_pct"), "traffic_source": traffic_src if d.get("traffic") else None}
        # "FSST HBM GB/s vs roofline" (BASELINE metric, second half): per-kernel figures are in `kernels`; end to end against the
        # fully fused bound of 180 B per sample
        fsst_ms = sum(per_kernel[k]["ms_per_step"] for k in FSST_KERNELS if k in per_kernel)
        fsst_rec = None
        if fsst_ms > 0:
            gbs = FSST_BOUND_BYTES * B * N_SAMPLES / (fsst_ms * 1e-3) / 1e9
            fsst_rec = {"ms_per_step": fsst_ms, "end_to_end_GB/s_vs_180B_bound": gbs, "frac_of_hbm_peak": gbs / peaks["hbm_gbs"],
                        "kernels": [k for k in FSST_KERNELS if k in per_kernel]}
        cpu = None
        configs = None
        if world == 1 and not args.no_cpu_baseline:
            ref = CpuReference(B)
            xs = shard_windows(all_windows, 0, B)
            ref.step(xs[:32])
            t0 = time.perf_counter()
            _, tf_, tl_ = ref.step(xs)
            dt = time.perf_counter() - t0
            cpu = {"value": B * N_SAMPLES / dt, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": "once: " + ref.describe(B),
                   "fsst_samples_per_s": B * N_SAMPLES / tf_, "lstm_samples_per_s": B * N_SAMPLES / tl_}
        if world == 1 and not args.no_side_configs:
            configs = side_configs(dev, flush, peaks, _lib)
        cm16 = [int(round(v)) for v in cfg4_host[:16].tolist()]
        line = {
            "metric": "PCG samples/s through FSST+BiLSTM", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (FSST fp32; LSTM gate GEMMs split-fp16 x3 on tcgen05, fp32 accumulate)"
            if os.environ.get("HSSB_LSTM_IMPL", "auto") != "simt" else "f32",
            "data": "synthetic",
            "config": bench_config(B),
            "timing": {"l2": "flushed between timed steps (256 MiB write)", "lstm_impl": os.environ.get("HSSB_LSTM_IMPL", "auto"),
                       "pipeline": "hss.pipeline.SegmentationPipeline: step i = BiLSTM of batch i + FSST of batch i+1 (a second stream, behind the model's side gate, "
                                   "on the SMs the recurrences leave idle); the first step also runs its own FSST, the last one prefetches nothing: K transforms "
                                   "and K model passes inside the K timed steps; results bit-identical to the sequential calls" if args.pipeline else "none (--no-pipeline)",
                       "collective": "one all-reduce of the 18-scalar metric state after the last step, inside the timed region",
                       "allreduce_ms": ms_allreduce,
                       "ms_per_step_without_pipeline": ms_sequential, "ms_of_each_step_without_pipeline": ms_sequential_steps,
                       "ms_of_each_step": [round(a.elapsed_time(b), 3) for a, b in ev[:-1]]},
            "e2e": {"value": samples / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": int(x_host.numel() * 4),
                    "d2h_bytes_per_step": int(lab_slots[0].numel() * 4 + 18 * 8), "pipeline": "double buffered: results of step i-1 read on the host while step i runs"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "kernels": per_kernel, "fsst": fsst_rec, "cpu_baseline": cpu,
            "configs": configs,
            "train_step_ms": configs["config3_training_step"]["ms"] if configs else None,
            "metrics_of_timed_steps": {"count": final_metrics["count"], "loss": final_metrics["loss"], "micro_accuracy": final_metrics["micro_accuracy"]},
            "config4_confusion": {"windows": CONFIG4_WINDOWS if cm16 and sum(cm16) else 0, "cm": cm16, "loss_sum": float(cfg4_host[16]),
                                  "crc32": zlib.crc32(",".join(map(str, cm16)).encode()),
                                  "note": "metric state over the same 4096 windows / one [2,4096,240] h0,c0 draw whatever N is: equal cm across N"},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--windows", type=int, default=WINDOWS_PER_GPU, help="windows per GPU")
    ap.add_argument("--ref-batch", type=int, default=0, help="windows per step of the reference arm (0 = the same as --windows)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-configs", action="store_true", help="skip the config 2 / 3 / 5 sub-records (N = 1 only)")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="do not prefetch the next step's FSST on a second stream under the current step's BiLSTM (hss.pipeline): every step "
                         "then runs its own transform first (0.5 ms per step slower at 512 windows)")
    args = ap.parse_args()
    args.pipeline = not args.no_pipeline
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29541", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
