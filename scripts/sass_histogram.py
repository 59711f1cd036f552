#!/usr/bin/env python
"""Opcode histogram of every kernel in libkzb200.so (cuobjdump -sass), written to profiles/sass_<kernel>.txt.

What proves a Blackwell-native kernel (B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG =
TMA loads / stores (UTMALDG...IM2COL = im2col mode), UTCBAR = tcgen05.commit, SYNCS = mbarrier.  No HMMA (legacy mma.sync) anywhere."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
lib = ROOT / "kzero_b200" / "libkzb200.so"
out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
kernels = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1
demangle = subprocess.run(["c++filt"] + list(kernels), capture_output=True, text=True).stdout.split("\n")
(ROOT / "profiles").mkdir(exist_ok=True)
summary = []
KEY = ("UTC", "LDTM", "STTM", "UTMA", "UBLKCP", "SYNCS", "HMMA", "UCGABAR", "ACQBULK", "CCTL")
for (mangled, hist), name in zip(kernels.items(), demangle):
    short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "")).split("::")[-1].replace("<", "_").replace(">", "").replace(" ", "")
    total = sum(hist.values())
    key = {op: n for op, n in hist.items() if op.startswith(KEY)}
    lines = [f"{name}", f"{total} SASS instructions ({total * 16 // 1024} KiB of code)", "", "-- tensor core / TMA / barrier opcodes"]
    lines += [f"{n:6d}  {op}" for op, n in sorted(key.items(), key=lambda kv: (-kv[1], kv[0]))]
    lines += ["", "-- full histogram"] + [f"{n:6d}  {op}" for op, n in sorted(hist.items(), key=lambda kv: (-kv[1], kv[0]))]
    (ROOT / "profiles" / f"sass_{short}.txt").write_text("\n".join(lines) + "\n")
    fam = collections.Counter()
    for op, n in key.items():
        fam[re.match(r"[A-Z]+", op).group(0)] += n
    summary.append(f"{short:28s} {total:6d} instr  " + "  ".join(f"{k} {v}" for k, v in sorted(fam.items())))
(ROOT / "profiles" / "sass_summary.txt").write_text("\n".join(summary) + "\n")
print("\n".join(summary))
