#!/usr/bin/env python
"""Development aid: per-SOURCE-LINE stall samples and instruction counts of one kernel from an ncu capture taken with
--import-source on.  ncu's CSV source page is per SASS instruction; the line table comes from `nvdisasm -g` of the same
library (instruction k of the function in both listings is the same instruction).
usage: python tools/ncu_lines.py REPORT.ncu-rep KERNEL_SUBSTRING [LIBRARY.so] [min_pct] [SUBSTRING_IN_THE_DISASSEMBLY]
(the two tools print template arguments differently: "(int)2, (int)32" in the report, "2, 32" in c++filt's output)"""
import csv, os, re, subprocess, sys, tempfile

rep, kern = sys.argv[1], sys.argv[2]
so = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cova_b200", "libcova_b200.so")
min_pct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
dis_kern = sys.argv[5] if len(sys.argv) > 5 else kern
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# several kernels may be in the report: take the first block whose name matches
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
blk = next(b for b in blocks if kern in b["name"])
hdr = blk["rows"][0]
iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
inst = [(int(r[iS] or 0), int(r[iI] or 0), r[hdr.index("Source")]) for r in blk["rows"][1:] if len(r) > iI and r[iS].isdigit()]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin") and "host_abi" not in f][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# locate the function: ".text.<mangled>" section whose demangled name contains kern
lines, on, cur_line, cur_file = [], False, None, None
want = None
for l in dis:
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        on = dis_kern in name and (want is None or want == name)
        if on:
            want = name
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_file, cur_line = os.path.basename(m.group(1)), int(m.group(2))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append((cur_file, cur_line))
n = min(len(lines), len(inst))
if len(lines) != len(inst):
    print(f"# warning: {len(inst)} instructions in the report, {len(lines)} in the disassembly", file=sys.stderr)
agg = {}
for (f, ln), (s, i, _) in zip(lines[:n], inst[:n]):
    a = agg.setdefault((f, ln), [0, 0])
    a[0] += s; a[1] += i
tot_s, tot_i = sum(a[0] for a in agg.values()) or 1, sum(a[1] for a in agg.values()) or 1
print(f"# {blk['name']}: {tot_s} samples, {tot_i} warp instructions")
src_cache = {}
for (f, ln), (s, i) in sorted(agg.items(), key=lambda kv: (kv[0][0] or "", kv[0][1] or 0)):
    if 100.0 * s / tot_s < min_pct and 100.0 * i / tot_i < min_pct:
        continue
    text = ""
    p = os.path.join(os.path.dirname(so), "csrc", f or "")
    if f and os.path.exists(p):
        src_cache.setdefault(p, open(p).read().splitlines())
        if ln and ln <= len(src_cache[p]):
            text = src_cache[p][ln - 1].strip()[:100]
    print(f"{100.0 * s / tot_s:5.1f}% samples {100.0 * i / tot_i:5.1f}% inst  {f}:{ln}  {text}")
