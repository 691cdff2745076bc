#!/usr/bin/env python
"""bench.py -- PCG samples/s through FSST + BiLSTM on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                          # the reference algorithm on host cores

One step = one pass of the hot path over one batch of synthetic PCG windows per GPU:
FSST (3 kernels) -> BiLSTM segmenter (eval forward) -> argmax labels -> 4x4 confusion counts
(-> NCCL all-reduce of the 16 counters when N > 1; no other collective).  Weak scaling: every rank
owns WINDOWS_PER_GPU windows (BASELINE config 4's shard: 4096 windows / 8 GPUs = 512 per GPU).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

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
# useful FLOPs (1x, the fp32 contraction) per PCG sample of every launch of the kernel in one step
FLOP_PER_SAMPLE = {"tc_inproj_l0": 2 * 44 * 1920.0, "tc_inproj_l1": 2 * 480 * 1920.0, "simt_inproj": 2 * (44 + 480) * 1920.0,
                   "tc_recurrent": 2 * 2 * 240 * 1920.0, "simt_recurrent": 2 * 2 * 240 * 1920.0}
BYTES_PER_SAMPLE = {"stft_hop1": 4 + 2 * K_BINS * 8, "if_reassign": 2 * K_BINS * 8 + KT * 8, "normalise": 16 * KT}


def workload_name(windows: int) -> str:
    return (f"config 4 shard: FSST(kaiser128,25-200Hz,stack)+BiLSTM(44->240x2x2->4), {windows} windows x {N_SAMPLES} "
            "samples per GPU")


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tensor_tflops": p["bf16_tflops_sustained"], "tensor_tflops_burst": p["bf16_tflops"],
                "source": "MEASURED_PEAKS.json (of measured)"}
    return {"hbm_gbs": 6650.0, "tensor_tflops": 1400.0, "tensor_tflops_burst": 1590.0, "source": "B200_PROFILING.md fallback (of fallback)"}


TRAFFIC_FILE = "r01d_traffic.json"


def traffic_source():
    return os.path.join("profiles", TRAFFIC_FILE)


def load_traffic() -> dict:
    path = os.path.join(ROOT, "profiles", TRAFFIC_FILE)
    if not os.path.exists(path):
        return {}
    try:
        return json.load(open(path))["kernels"]
    except (ValueError, KeyError):
        return {}


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
# synthetic workload (same generator as the parity tests)
# --------------------------------------------------------------------------------------------------
def make_windows(n_windows: int, seed: int) -> np.ndarray:
    from workloads import tiled_windows

    return tiled_windows(n_windows, N_SAMPLES, FS, seed)


def synthetic_targets(n_windows: int) -> np.ndarray:
    from workloads import synthetic_targets as st

    return st(n_windows, N_SAMPLES)


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
        self.window = fo.reference_window(NWIN)
        self.dwindow = fo.dtwin(self.window, FS)
        self.params, self.h0, self.c0 = lo.reference_params(68, 2 * KT, batch, 240)
        self.fsst_threads = int(self.lib.hsso_num_threads())
        self.torch_threads = torch.get_num_threads()

    def step(self, x: np.ndarray):
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


def run_reference(args, rank: int):
    if rank != 0:
        return
    batch = args.ref_batch
    ref = CpuReference(batch)
    x = make_windows(batch, 68)
    for _ in range(args.warmup):
        ref.step(x)
    t_f = t_l = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, a, b = ref.step(x)
        t_f += a; t_l += b
    dt = time.perf_counter() - t0
    value = batch * N_SAMPLES * args.steps / dt
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": "PCG samples/s through FSST+BiLSTM", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.windows), "windows_per_gpu": args.windows, "samples_per_window": N_SAMPLES,
                   "sample_windows_per_step": batch},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{batch} windows x {N_SAMPLES} samples per step; FSST = C/OpenMP restatement of MATLAB fsst "
                                   f"({ref.fsst_threads} threads; libssq itself is unobtainable), BiLSTM = torch-CPU nn.LSTM restatement of "
                                   f"segmenter.py ({ref.torch_threads} threads)",
                         "fsst_samples_per_s": batch * N_SAMPLES * args.steps / t_f, "lstm_samples_per_s": batch * N_SAMPLES * args.steps / t_l},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------------------
def run_ours(args, rank: int, local_rank: int, world: int):
    import torch.distributed as dist
    from hss import _lib
    from hss.model.segmenter import HeartSoundSegmenter
    from hss.sharding import allreduce_counts, confusion_counts
    from hss.transforms import FSST
    from workloads import reference_window

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.windows
    peaks = load_peaks()

    x_host = torch.from_numpy(make_windows(B, 68 + 1000 * rank)).pin_memory()
    y_dev = torch.from_numpy(synthetic_targets(B)).to(dev)
    x_dev = x_host.to(dev)
    fsst = FSST(FS, window=reference_window(NWIN), truncate_freq=BAND, stack=True)
    torch.manual_seed(68)
    model = HeartSoundSegmenter(input_size=2 * KT, batch_size=B).eval()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step(x):
        feats = fsst.batch(x)
        labels = model.predict(feats)
        cm = confusion_counts(labels, y_dev)
        return labels, allreduce_counts(cm)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step(x_dev)
    barrier()

    # ---- device-resident timing: K steps, CUDA events per step, L2 flushed between steps ----
    sampler.mark()
    _lib.prof_enable(True)
    _lib.prof_read()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    gc.collect()        # (before the barrier: a gen-2 collection takes 10-50 ms and differs per rank -- after it, the skew would
    gc.disable()        #  sit in the first step's all-reduce)
    barrier()
    for a, b in ev:
        flush.zero_()
        a.record()
        step(x_dev)
        b.record()
    barrier()
    gc.enable()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    prof = _lib.prof_read()
    _lib.prof_enable(False)
    clocks = sampler.stop()

    # ---- end to end: pinned host input -> H2D -> path -> labels + counters back on the host ----
    # (the input lands in a preallocated device buffer and the Python GC is off inside the timed region: a torch allocator miss
    #  or a gen-2 collection costs 30-90 ms of host time, which used to hit one of the five steps every few runs)
    # Double buffered like a production ingest loop: step i is enqueued (H2D copy, kernels, D2H copies into pinned slot i % 2),
    # then the host waits for step i-1's event and reads its results -- every step's labels and counters reach the host inside
    # the timed region, one step behind the device, so host launch latency and scheduling jitter overlap with device work.
    lab_slots = [torch.empty((B, N_SAMPLES), dtype=torch.int32).pin_memory() for _ in range(2)]
    cm_slots = [torch.empty((4, 4), dtype=torch.int64).pin_memory() for _ in range(2)]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    x_in = torch.empty_like(x_dev)

    def e2e_enqueue(i):
        x_in.copy_(x_host, non_blocking=True)
        labels, cm = step(x_in)
        lab_slots[i & 1].copy_(labels, non_blocking=True)
        cm_slots[i & 1].copy_(cm, non_blocking=True)
        done[i & 1].record()

    def e2e_collect(i):
        done[i & 1].synchronize()
        return int(cm_slots[i & 1].sum()) + int(lab_slots[i & 1][0, 0])      # the host reads the step's results

    def e2e_run(n):
        seen = 0
        for i in range(n):
            e2e_enqueue(i)
            if i:
                seen += e2e_collect(i - 1)
            wall.append(time.perf_counter())
        return seen + e2e_collect(n - 1)

    wall = []
    e2e_run(max(args.warmup, 3))
    gc.collect()
    gc.disable()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    debug = bool(os.environ.get("HSSB_BENCH_DEBUG"))
    wall = [time.perf_counter()]
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    gc.enable()
    ms_e2e = e0.elapsed_time(e1)
    cm_host = cm_slots[(args.steps - 1) & 1].clone()
    if debug:
        print(f"rank {rank} e2e wall per step (ms):", [round(1e3 * (b - a), 3) for a, b in zip(wall, wall[1:])], "events total", ms_e2e, file=sys.stderr)

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    samples = B * N_SAMPLES * world * args.steps
    value = samples / (ms * 1e-3)

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
            elif name in FLOP_PER_SAMPLE:
                flop = FLOP_PER_SAMPLE[name]
                if name == "tc_recurrent" and "tc_inproj_l0" not in prof:
                    flop += FLOP_PER_SAMPLE["tc_inproj_l0"]        # layer 1's input projection is fused into the recurrence kernel
                tf = flop * units / (tot / args.steps * 1e-3) / 1e12
                entry.update({"bound": "tensor", "achieved": tf, "peak": peaks["tensor_tflops"], "unit": "TFLOP/s", "frac": tf / peaks["tensor_tflops"],
                              "note": "useful FLOPs (1x) over all launches of this kernel in a step"})
            per_kernel[name] = entry
        # ncu evidence committed under profiles/ (scripts/gpu_profiles.sh + scripts/ncu_traffic.py): DRAM bytes per launch and
        # tensor-pipe activity of every kernel of this workload; only attached when the workload is the profiled one
        traffic = load_traffic() if (B == WINDOWS_PER_GPU) else {}
        for name, entry in per_kernel.items():
            t = traffic.get(name)
            if t:
                entry["traffic"] = t["dram_bytes_per_launch"]
                if t["tensor_pipe_active_pct"] > 0:
                    entry["ncu_tensor_pipe_active_pct"] = t["tensor_pipe_active_pct"]
        dominant = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_step"])
        d = per_kernel[dominant]
        roofline = {"kernel": dominant, "bound": d.get("bound"), "achieved": d.get("achieved"), "peak": d.get("peak"), "unit": d.get("unit"),
                    "frac": d.get("frac"), "traffic": d.get("traffic"), "share_of_step": d["ms_per_step"] / (ms / args.steps),
                    "peak_source": peaks["source"],
                    "note": "achieved = useful fp32-equivalent FLOPs (1x) of the W_hh contraction / launch time; the kernel issues 3 fp16 MMAs "
                            "per product on 8-CTA clusters (96 of 148 SMs at 512 windows), ncu sm__pipe_tensor_cycles_active in "
                            "ncu_tensor_pipe_active_pct" if dominant == "tc_recurrent" else None,
                    "ncu_tensor_pipe_active_pct": d.get("ncu_tensor_pipe_active_pct"), "traffic_source": traffic_source() if d.get("traffic") else None}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            ref = CpuReference(50)
            xs = make_windows(50, 68)
            ref.step(xs)
            t0 = time.perf_counter()
            _, tf_, tl_ = ref.step(xs)
            dt = time.perf_counter() - t0
            cpu = {"value": 50 * N_SAMPLES / dt, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": f"50 windows x {N_SAMPLES} samples once (FSST C/OpenMP restatement {ref.fsst_threads} thr + torch-CPU BiLSTM {ref.torch_threads} thr)",
                   "fsst_samples_per_s": 50 * N_SAMPLES / tf_, "lstm_samples_per_s": 50 * N_SAMPLES / tl_}
        line = {
            "metric": "PCG samples/s through FSST+BiLSTM", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (FSST fp32; LSTM gate GEMMs split-fp16 x3 on tcgen05, fp32 accumulate)"
            if os.environ.get("HSSB_LSTM_IMPL", "auto") != "simt" else "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(B),
                       "windows_per_gpu": B, "samples_per_window": N_SAMPLES, "l2": "flushed between timed steps (256 MiB write)",
                       "lstm_impl": os.environ.get("HSSB_LSTM_IMPL", "auto")},
            "e2e": {"value": samples / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": int(x_host.numel() * 4),
                    "d2h_bytes_per_step": int(lab_slots[0].numel() * 4 + 128), "pipeline": "double buffered: results of step i-1 read on the host while step i runs"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "kernels": per_kernel, "cpu_baseline": cpu,
            "confusion_total": int(cm_host.sum()),
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
    ap.add_argument("--ref-batch", type=int, default=50, help="windows per step of the reference arm (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
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
