"""Aggregate an `ncu --page source --csv` export (SASS rows with stall samples) by source line / function,
using nvdisasm line info of the same library.  usage: ncu_source_hist.py source.csv [lib.so]"""
import collections, csv, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src_csv = sys.argv[1]
so = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "sesameai-tts_b200", "lib", "libcsm_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.startswith("api.") and f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
if not sass:
    sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
# offset -> (file:line chain)
inside = False; cur = "?"; off2line = {}; fresh = True
for l in sass.splitlines():
    if l.startswith(".text."):
        inside = "k_frame_mega" in l; continue
    if not inside: continue
    m = re.search(r'//## File "(.*?)", line (\d+)(?: inlined at "(.*?)", line (\d+))?', l)
    if m:
        if fresh:  # the first entry of a chain is the innermost frame
            cur = f"{os.path.basename(m.group(1))}:{m.group(2)}"
            fresh = False
        continue
    m = re.match(r"^\s+/\*([0-9a-f]+)\*/\s+([A-Z@].*?);", l)
    if m:
        off2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
        fresh = True
rows = list(csv.reader(l for l in open(src_csv) if l.startswith('"')))
hdr = rows[1]
ia, isamp, isrc = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Source")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = None
by_line = collections.Counter(); by_reason = collections.Counter(); line_reason = collections.defaultdict(collections.Counter)
line_inst = collections.Counter(); iexec = hdr.index("Instructions Executed")
tot = 0
for r in rows[2:]:
    if len(r) != len(hdr): continue
    a = int(r[ia], 16)
    if base is None: base = a
    line, ins = off2line.get(a - base, ("?", r[isrc]))
    n = int(r[isamp] or 0); tot += n
    by_line[line] += n
    line_inst[line] += int(r[iexec] or 0)
    for i in stall_cols:
        v = int(r[i] or 0)
        if v: by_reason[hdr[i]] += v; line_reason[line][hdr[i]] += v
print("total samples", tot)
print("by stall reason:", ", ".join(f"{k[6:]} {v*100/tot:.1f}%" for k, v in by_reason.most_common(12)))
srcs = {}
def text(line):
    f, ln = line.split(":") if ":" in line else (line, "0")
    p = os.path.join(ROOT, "sesameai-tts_b200/csrc", f)
    if os.path.exists(p):
        if p not in srcs: srcs[p] = open(p).read().splitlines()
        i = int(ln) - 1
        if 0 <= i < len(srcs[p]): return srcs[p][i].strip()[:80]
    return ""
print(f"{'samples':>8s} {'%':>6s} {'winst':>10s}  line / top stall reasons / text")
for line, n in by_line.most_common(45):
    rs = ", ".join(f"{k[6:]} {v*100//max(n,1)}%" for k, v in line_reason[line].most_common(3))
    print(f"{n:8d} {n*100/tot:6.2f} {line_inst[line]:10d}  {line:22s} [{rs}]  {text(line)}")
