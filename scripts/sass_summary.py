#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove Blackwell-native code (runs here, no GPU needed):

    python scripts/sass_summary.py [lib] > profiles/r02_sass_summary.txt

tcgen05.mma -> UTC*MMA, tcgen05.ld / st -> LDTM / STTM, TMA tensor + bulk copies -> UTMALDG / UTMASTG / UBLKCP,
tcgen05.commit -> UTCBAR, mbarrier -> SYNCS, cluster barrier -> UCGABAR, legacy mma.sync -> HMMA.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "heart-sounds-segmentation_b200", "lib", "libhssb.so")
PATTERNS = [("UTCHMMA", r"\bUTCHMMA"), ("UTCHMMA.2CTA", r"UTCHMMA\.2CTA"), ("UTCHMMA A_KEEP/A_REUSE", r"UTCHMMA.*\.A_(KEEP|REUSE)"),
            ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTCBAR", r"\bUTCBAR"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
            ("UBLKCP", r"\bUBLKCP"), ("UBLKCP.MULTICAST", r"UBLKCP.*MULTICAST"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("UCGABAR (cluster barrier)", r"\bUCGABAR"),
            ("MUFU", r"\bMUFU"), ("HMMA (legacy mma.sync)", r"\bHMMA")]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = {}
names = re.findall(r"Function : (\S+)", sass)
if names:
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    demangle = dict(zip(names, out))
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = demangle.get(m.group(1), m.group(1))
        cur = re.sub(r"\((?:int|bool)\)", "", cur)                      # template arguments without the casts
        cur = re.sub(r"\([^()]*\)$", "", cur).replace("void ", "").replace("hssb::", "").replace("<unnamed>::", "")
        counts[cur] = collections.Counter()
        continue
    if cur is None or "/*" not in line:
        continue
    counts[cur]["instructions"] += 1
    for label, pat in PATTERNS:
        if re.search(pat, line):
            counts[cur][label] += 1
print(f"# SASS summary of {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a): instruction counts per kernel")
tot = collections.Counter()
for k, c in counts.items():
    hits = ", ".join(f"{label} {c[label]}" for label, _ in PATTERNS if c[label])
    print(f"{k:70s} {c['instructions']:6d} instr   {hits}")
    tot.update(c)
print("\nTOTAL  " + ", ".join(f"{label} {tot[label]}" for label, _ in PATTERNS if tot[label]))
