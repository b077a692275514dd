"""Joins an ncu `--page source --csv` dump (per-SASS-instruction executed counts and stall samples)
with nvdisasm line info, and prints instructions executed / stall samples per CUDA source line.

usage: python tools/ncu_by_line.py <report.ncu-rep> <kernel-mangled-substring> [lib.so]
"""
import csv
import collections
import os
import re
import subprocess
import sys
import tempfile

rep, kern = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else "rllib_warehouse_b200/lib/libwh_b200.so"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# find section
start = next(i for i, l in enumerate(dis) if l.startswith("\t.section\t.text.") and kern in l)
lines_by_off, cur = {}, ("?", 0)
inline_stack = None
for l in dis[start + 1:]:
    if l.startswith("\t.section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        lines_by_off[int(m.group(1), 16)] = (cur, m.group(2))
csv_txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(csv_txt.splitlines()))
hdr = rows[1]
ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = int(rows[2][ia], 16)
agg = collections.defaultdict(lambda: [0, 0, 0])
tot_e = tot_s = 0
for r in rows[2:]:
    if len(r) <= ie or not r[ia].startswith("0x"):
        break
    off = int(r[ia], 16) - base
    (f, ln), sass = lines_by_off.get(off, (("?", 0), ""))
    a = agg[(f, ln)]
    a[0] += int(r[ie]); a[1] += int(r[isamp]); a[2] += 1
    tot_e += int(r[ie]); tot_s += int(r[isamp])
print(f"total warp-instructions {tot_e}, samples {tot_s}")
src_cache = {}
def src(f, ln):
    for d in ("rllib_warehouse_b200/csrc", "include"):
        p = os.path.join(d, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            return src_cache[p][ln - 1].strip()[:90] if 0 < ln <= len(src_cache[p]) else ""
    return ""
print(f"{'file:line':28s} {'inst%':>6s} {'samp%':>6s} {'sass':>5s}  source")
for (f, ln), (e, s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get("TOP", "45"))]:
    print(f"{f+':'+str(ln):28s} {100*e/tot_e:6.2f} {100*s/max(tot_s,1):6.2f} {n:5d}  {src(f, ln)}")
