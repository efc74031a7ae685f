"""Attribute an `ncu --page source --csv` SASS export to CUDA source lines.

    python profiles/line_profile.py <src.csv> <cubin> <kernel-name-substring> [top]

ncu's CSV has no line numbers; `nvdisasm --print-line-info` lists the same instructions in the same order
with `//## File "...", line N` markers, so the two are zipped by position.
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

src_csv, cubin, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
text = subprocess.run(["nvdisasm", "--print-line-info", cubin], stdout=subprocess.PIPE, text=True).stdout
lines_of = []       # (file, line) per instruction of the kernel, in order
cur, inside = ("?", 0), False
for ln in text.splitlines():
    if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln):
        inside = kname in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
        lines_of.append(cur)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ii, wi = hdr.index('Instructions Executed'), hdr.index('# Samples')
data = [(int(r[ii]), int(r[wi])) for r in rows[2:] if len(r) > wi and r[ii].isdigit()]
# the csv may hold several launches of the kernel back to back
n = len(lines_of)
assert n and len(data) % n == 0, (len(data), n)
agg = defaultdict(lambda: [0, 0])
for k, (ins, smp) in enumerate(data):
    key = lines_of[k % n]
    agg[key][0] += ins
    agg[key][1] += smp
ti = sum(v[0] for v in agg.values())
ts = sum(v[1] for v in agg.values())
srcs = {}
print("instructions %d  samples %d  (%d SASS lines, %d launches)" % (ti, ts, n, len(data) // n))
for (f, l), (ins, smp) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in srcs:
        try:
            srcs[f] = open("/root/repo/retinanet-for-table-detection_b200/csrc/" + f).read().splitlines()
        except Exception:
            srcs[f] = []
    code = srcs[f][l - 1].strip()[:100] if 0 < l <= len(srcs[f]) else ""
    print("%5.1f%% instr %5.1f%% stall  %s:%d  %s" % (100.0 * ins / ti, 100.0 * smp / max(1, ts), f, l, code))
