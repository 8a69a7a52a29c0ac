#!/usr/bin/env python
"""Executed warp-instructions of one kernel of an .ncu-rep aggregated by CUDA source line.
usage: ncu_by_line.py <rep> <kernel-regex> <object.o> <mangled-substring> <source.cu> [top]
Joins the SASS page of the report (instruction order) with nvdisasm --print-line-info of the object."""
import csv, io, re, subprocess, sys, collections, tempfile, os, glob
rep, rx, obj, mangled, srcfile = sys.argv[1:6]
top = int(sys.argv[6]) if len(sys.argv) > 6 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
dis = subprocess.run(["nvdisasm", "--print-line-info", "-c"] + glob.glob(tmp + "/*.cubin"), capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and mangled in l)
cur, seq = None, []
for l in dis[start + 1:]:
    if l.startswith("//---------------------"):
        break
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        cur = int(m.group(2)) if m.group(1).endswith(os.path.basename(srcfile)) else -1
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append((cur, m.group(2)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[h]; ii = hdr.index("Instructions Executed"); si = hdr.index("# Samples")
data = [(r[1].strip(), int(r[ii]), int(r[si])) for r in rows[h + 1:] if len(r) > ii and r[ii].isdigit()]
print(f"sass in object {len(seq)}  sass in report {len(data)}")
by, bys = collections.Counter(), collections.Counter()
for (ln, _s), (_src, n, sm) in zip(seq, data):
    by[ln] += n; bys[ln] += sm
tot, ts = sum(by.values()), sum(bys.values())
src = open(srcfile).read().split("\n")
for ln, n in sorted(by.items(), key=lambda t: -t[1])[:top]:
    text = src[ln - 1].strip()[:95] if ln and ln > 0 else "(other file)"
    print(f"{ln:5d} {n / tot * 100:5.1f}% smp {bys[ln] / ts * 100:5.1f}%  {text}")
