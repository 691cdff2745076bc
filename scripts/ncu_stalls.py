#!/usr/bin/env python
"""Warp-stall sampling of an ncu report by code region (run here, no GPU needed).

    python scripts/ncu_stalls.py gpurun_out/x.ncu-rep [kernel-name-substring] > profiles/x_stalls.txt

For every captured launch whose name contains the substring: the share of warp samples between consecutive
"landmark" instructions (barrier waits, TMEM loads, MUFU blocks, bulk copies, TMA stores, global loads / stores)
with the three dominant stall reasons -- the per-instruction source page of ncu, condensed.
"""
import csv
import io
import subprocess
import sys

MARKS = ["LDTM", "SYNCS", "UBLKCP", "BAR.SYNC", "UTMASTG", "LDG", "STG", "MUFU.EX2", "MUFU.RCP", "STS", "UTCHMMA", "NANOSLEEP", "WARPSYNC",
         "FENCE", "MEMBAR", "UTMACMDFLUSH", "DEPBAR", "UTMALDG", "SHFL", "STTM"]
ALWAYS = ("SYNCS", "UBLKCP", "BAR.SYNC", "LDTM", "UTMASTG")

path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(raw)):
    if r and r[0] == "Kernel Name":
        cur = []
        blocks.append((r[1], cur))
    elif cur is not None:
        cur.append(r)
seen = set()
for name, b in blocks:
    if want not in name or not b:
        continue
    hdr, data = b[0], b[1:]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[isamp] or 0) for r in data)
    if (name, tot) in seen:
        continue
    seen.add((name, tot))
    print(f"=== {name[:110]}\n    {tot} warp samples, {len(data)} SASS instructions")
    acc, accst, last, n = 0, {}, None, 0
    for r in data:
        s = int(r[isamp] or 0)
        acc += s
        n += 1
        for i in cols:
            v = int(r[i] or 0)
            if v:
                accst[hdr[i][6:]] = accst.get(hdr[i][6:], 0) + v
        m = [x for x in MARKS if x in r[isrc]]
        if m and (m[0] != last or m[0] in ALWAYS):
            if acc * 200 >= tot:        # >= 0.5 % of the samples
                top = ", ".join(f"{k} {100 * v / max(acc, 1):.0f}%" for k, v in sorted(accst.items(), key=lambda kv: -kv[1])[:3])
                print(f"    {100 * acc / tot:5.1f}%  {n:4d} instr up to  {r[isrc].strip()[:52]:52s} executed {r[iex]:>9s}   [{top}]")
            acc, accst, last, n = 0, {}, m[0], 0
