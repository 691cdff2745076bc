#!/usr/bin/env python
"""Summarise an ncu report (run here, no GPU needed): one block per captured launch with the metrics the
roofline discussion uses.   python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(f"== {r[idx['Kernel Name']][:100]}   (ID {r[idx['ID']]})")
        for key, label in WANT:
            if key in idx:
                print(f"   {label:26s} {r[idx[key]]:>16s} {units[idx[key]]}")
        rd, wr, dur = (r[idx.get(k, 0)] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
        try:
            mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}
            b = float(rd) * mult[units[idx["dram__bytes_read.sum"]]] + float(wr) * mult[units[idx["dram__bytes_write.sum"]]]
            t = float(dur) * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}[units[idx["gpu__time_duration.sum"]]]
            print(f"   {'dram traffic':26s} {b / 1e6:16.3f} MB  -> {b / t / 1e9:.1f} GB/s under ncu (cold, serialised)")
        except (ValueError, KeyError):
            pass


if __name__ == "__main__":
    main(sys.argv[1])
