"""Per-phase timeline of the decode megakernel (CTA 0, %globaltimer): where a frame's time goes."""
import collections
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sesameai-tts_b200"))
sys.path.insert(0, ROOT)
import torch

import bench
from sesameai import _native, synthetic as syn

model = bench.build_product(torch.device("cuda", 0), 1)
L = _native.lib()
nph = L.csm_debug_set_trace(model._ctx, None)
NCTA = torch.cuda.get_device_properties(0).multi_processor_count
buf = torch.zeros(NCTA, nph, 16, dtype=torch.int64, device="cuda")
tok, msk, pos = syn.text_prompt(1, 32, 4321, device="cuda")
model.reset_caches()
s = model.generate_frame(tok, msk, pos, 1.0, 1)
t = torch.zeros(1, 1, 33, dtype=torch.int64, device="cuda")
m = torch.ones(1, 1, 33, dtype=torch.bool, device="cuda"); m[..., -1] = False
p = torch.full((1, 1), 31, dtype=torch.int64, device="cuda")
for i in range(4):
    t[:, 0, :32] = s; p += 1
    if i == 3:
        L.csm_debug_set_trace(model._ctx, buf.data_ptr())
    s = model.generate_frame(t, m, p, 1.0, 1)
torch.cuda.synchronize()
L.csm_debug_set_trace(model._ctx, None)
tr = buf.cpu().numpy()
# phase kinds in table order (api.cu build_mega_phases)
kinds = ["embed"]
for l in range(16):
    kinds += ["bb.qkv", "bb.attn", "bb.o", "bb.gu", "bb.down"]
kinds += ["c0head", "sample"]
qkv_table = nph == 641 - 30  # the first layer's q / k / v of steps >= 2 come from the table gather of the sample phase
for i in range(1, 32):
    for l in range(4):
        kinds += (["d.o+attn", "d.gu", "d.down"] if (qkv_table and i >= 2 and l == 0) else ["d.qkv", "d.o+attn", "d.gu", "d.down"])
    kinds += ["d.head", "sample"]
assert len(kinds) == nph, (len(kinds), nph)
import numpy as np
tr = tr.astype(np.int64)                       # [ncta, nph, 4]: start, inputs staged, partial sums done, end
t0 = tr[:, 0, 0].min()
start, staged, mid, end = (tr[:, :, k] for k in range(4))
ck = tr[:, :, 4:16]
active = staged > 0                            # CTAs that had work in the phase
done = np.where(active, end, 0).max(axis=0)    # phase completion = last active CTA's end
done = np.where(active.any(axis=0), done, end.max(axis=0))
prev = np.concatenate([[t0], done[:-1]])
big = np.iinfo(np.int64).max
agg = collections.defaultdict(lambda: [0] + [0.0] * 6)
for i, k in enumerate(kinds):
    a = agg[k]
    a[0] += 1
    a[1] += (done[i] - prev[i]) / 1e3                                   # critical path of the phase
    if active[:, i].any():
        st = np.where(active[:, i], staged[:, i], big)
        a[2] += (st.min() - prev[i]) / 1e3                              # hand-off: last producer end -> FIRST consumer has its inputs
        a[3] += (np.where(active[:, i], staged[:, i], 0).max() - prev[i]) / 1e3   # ... -> LAST consumer has its inputs
        a[4] += np.where(active[:, i], mid[:, i] - staged[:, i], 0).max() / 1e3   # slowest weight stream + mma
        a[5] += np.where(active[:, i], end[:, i] - mid[:, i], 0).max() / 1e3      # slowest epilogue
        a[6] += active[:, i].sum()
print(f"frame total {(done[-1]-t0)/1e3:.1f} us over {nph} phases (all CTAs)")
print(f"{'phase':10s} {'n':>4s} {'crit us':>8s} {'1st in':>8s} {'last in':>8s} {'stream':>8s} {'epi':>8s} {'ctas':>6s}   (avg per phase)")
for k, a in agg.items():
    n = a[0]
    print(f"{k:10s} {n:4d} {a[1]/n:8.2f} {a[2]/n:8.2f} {a[3]/n:8.2f} {a[4]/n:8.2f} {a[5]/n:8.2f} {a[6]/n:6.0f}   total {a[1]:8.1f}")

# clock64 marks of thread 0 (cycles since phase start), averaged over the CTAs that had work
names = ["item+pre", "polled", "sumsq out", "inv ready", "idx math", "chunk wait", "mma done", "psum out", "csync", "epilogue", "refill+end"]
print("cycles since phase start at each mark (thread 0, active CTAs):")
print(f"{'phase':10s} " + " ".join(f"{n:>10s}" for n in names))
for kind in ["bb.qkv", "bb.o", "bb.gu", "bb.down", "d.qkv", "d.o+attn", "d.gu", "d.down", "d.head"]:
    idx = [i for i, k in enumerate(kinds) if k == kind]
    sel = active[:, idx]
    row = []
    for m in range(1, 12):
        d = (ck[:, idx, m] - ck[:, idx, 0])
        ok = sel & (ck[:, idx, m] > 0)
        row.append(d[ok].mean() if ok.any() else float("nan"))
    print(f"{kind:10s} " + " ".join(f"{v:10.0f}" for v in row))
