#!/usr/bin/env python
"""Per-kernel SASS evidence for profiles/: counts of the Blackwell mnemonics (B200_PROFILING.md, "What proves a
Blackwell-native kernel") in every kernel of the shipped library.
usage: python tools/sass_summary.py [cova_b200/libcova_b200.so] > profiles/r2_sass_summary.txt"""
import os, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cova_b200", "libcova_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "FFMA2", "HMMA", "LDG", "STG", "LDS", "STS", "ATOMS", "FENCE"]
kernels, cur = {}, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kernels[cur] = {k: 0 for k in MN}
        kernels[cur]["total"] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if cur and m:
        kernels[cur]["total"] += 1
        op = m.group(1)
        for k in MN:
            if op.startswith(k):
                kernels[cur][k] += 1
print(f"# {os.path.basename(so)}: SASS mnemonic counts per kernel (cuobjdump -sass); UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,")
print("# UBLKCP = cp.async.bulk (the bulk-copy engine; operands are 1-D strips, no tensor map: UTMALDG = 0), SYNCS = mbarrier ops")
print(f"{'kernel':100s} " + " ".join(f"{k:>8s}" for k in ["total"] + MN))
tot = {k: 0 for k in ["total"] + MN}
for name, c in sorted(kernels.items()):
    short = re.sub(r"\(.*", "", name).replace("cova::", "")
    print(f"{short[:100]:100s} " + " ".join(f"{c[k]:8d}" for k in ["total"] + MN))
    for k in tot:
        tot[k] += c[k]
print(f"{'ALL KERNELS':100s} " + " ".join(f"{tot[k]:8d}" for k in ["total"] + MN))
