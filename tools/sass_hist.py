"""Instruction-count histogram of k_frame_mega by source function (needs -lineinfo): which code is hot-path sized."""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "sesameai-tts_b200", "lib", "libcsm_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.startswith("api.") and f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
# function line ranges of mega.cuh / lm_kernels.cuh
def ranges(path):
    out, cur = [], None
    for i, l in enumerate(open(path), 1):
        m = re.match(r"^(?:template.*\n)?(?:__device__|__global__).*?\b(\w+)\s*\(", l)
        if m and not l.startswith(" "):
            out.append((i, m.group(1)))
    return out
files = {"mega.cuh": ranges(os.path.join(ROOT, "sesameai-tts_b200/csrc/mega.cuh")),
         "lm_kernels.cuh": ranges(os.path.join(ROOT, "sesameai-tts_b200/csrc/lm_kernels.cuh")),
         "common.cuh": ranges(os.path.join(ROOT, "sesameai-tts_b200/csrc/common.cuh"))}
def fn_of(f, ln):
    r = files.get(f)
    if not r: return f
    name = f
    for start, n in r:
        if start <= ln: name = n
        else: break
    return name
inside = False
cnt = collections.Counter(); per_line = collections.Counter(); cur = "?"; tot = 0
stack = "?"
for l in sass.splitlines():
    if l.startswith(".text."):
        inside = "k_frame_mega" in l
        continue
    if not inside: continue
    m = re.search(r'//## File "(.*?)", line (\d+)(.*)', l)
    if m:
        f = os.path.basename(m.group(1)); ln = int(m.group(2))
        cur = fn_of(f, ln)
        # outermost mega.cuh frame if this is inlined ("inlined at" info is on following lines; keep simple)
        continue
    if re.match(r"^\s+/\*[0-9a-f]+\*/\s+[A-Z@]", l):
        cnt[cur] += 1; tot += 1
print("total", tot, "instructions =", tot * 16 // 1024, "KB")
for k, v in cnt.most_common(40): print(f"{v:6d}  {k}")
