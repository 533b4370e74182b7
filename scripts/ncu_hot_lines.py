#!/usr/bin/env python
"""Attribute ncu warp-stall samples (SASS source page) to CUDA source lines via nvdisasm line info.

  python scripts/ncu_hot_lines.py REPORT.ncu-rep KERNEL_REGEX CUBIN MANGLED_SUBSTR [min_frac]
"""
import csv, io, re, subprocess, sys, collections

rep, kre, cubin, fsub = sys.argv[1:5]
minf = float(sys.argv[5]) if len(sys.argv) > 5 else 0.01
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
sass = [r for r in rows[2:] if len(r) > ci["# Samples"]]
# nvdisasm with line info
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
lines_of = []   # per instruction in function order: (file:line)
cur = None; infn = False
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+)", l)
    if m:
        infn = fsub in m.group(1); continue
    if not infn: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)), m.group(3)); continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        lines_of.append(cur)
if len(lines_of) != len(sass):
    print("warning: %d SASS rows in report vs %d in cubin" % (len(sass), len(lines_of)))
agg = collections.Counter(); insts = collections.Counter(); stall = collections.defaultdict(collections.Counter)
tot = 0
for r, ln in zip(sass, lines_of):
    s = int(r[ci["# Samples"]] or 0); tot += s
    key = (ln[0], ln[1]) if ln else ("?", 0)
    agg[key] += s; insts[key] += int(r[ci["Instructions Executed"]] or 0)
    for h in hdr:
        if h.startswith("stall_") and r[ci[h]] not in ("", "0"):
            stall[key][h[6:]] += int(r[ci[h]])
print("total samples", tot, "instructions", sum(insts.values()))
for key, s in agg.most_common():
    if s < tot * minf: break
    top = ", ".join("%s %d" % kv for kv in stall[key].most_common(3))
    print("%5.1f%%  %s:%d  inst %d  [%s]" % (100.0 * s / tot, key[0], key[1], insts[key], top))
