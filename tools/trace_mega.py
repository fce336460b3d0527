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
buf = torch.zeros(nph, 8, dtype=torch.int64, device="cuda")
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
for i in range(1, 32):
    for l in range(4):
        kinds += ["d.qkv", "d.o+attn", "d.gu", "d.down"]
    kinds += ["d.head", "sample"]
assert len(kinds) == nph, (len(kinds), nph)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
fine = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
for i, k in enumerate(kinds[:-1]):
    a = agg[k]
    a[0] += 1
    a[1] += (tr[i, 1] - tr[i, 0]) / 1e3   # work
    a[2] += (tr[i, 2] - tr[i, 1]) / 1e3   # CTA-local sync
    a[3] += (tr[i, 3] - tr[i, 2]) / 1e3   # grid barrier (arrive + wait for slowest CTA)
    f = fine[k]
    f[0] += 1
    for j in range(4):
        if tr[i, 4 + j] > 0:
            f[1 + j] += (tr[i, 4 + j] - tr[i, 0]) / 1e3
print(f"frame total {(tr[-1,1]-tr[0,0])/1e3:.1f} us over {nph} phases")
print(f"{'phase':10s} {'n':>4s} {'work us':>9s} {'csync us':>9s} {'grid us':>9s}   (avg per phase)")
tot = [0, 0, 0]
for k, a in agg.items():
    print(f"{k:10s} {a[0]:4d} {a[1]/a[0]:9.2f} {a[2]/a[0]:9.2f} {a[3]/a[0]:9.2f}   total {sum(a[1:]):8.1f}")
    for j in range(3): tot[j] += a[1 + j]
print("totals us: work %.1f csync %.1f grid %.1f" % tuple(tot))
print("fine marks (avg us since phase start): gemv: x staged / first chunk landed / first chunk done / refill issued; sample: start / sampled")
for k, f in fine.items():
    print(f"{k:10s} " + " ".join(f"{f[1+j]/f[0]:8.2f}" for j in range(4)))
