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
buf = torch.zeros(nph, 4, dtype=torch.int64, device="cuda")
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
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
for i, k in enumerate(kinds):
    a = agg[k]
    a[0] += 1
    nxt = tr[i + 1, 0] if i + 1 < nph else tr[i, 3]
    a[1] += (nxt - tr[i, 0]) / 1e3                                    # whole phase (start -> next start)
    if tr[i, 1] > 0: a[2] += (tr[i, 1] - tr[i, 0]) / 1e3              # inputs staged (hand-off wait + norm / attention)
    if tr[i, 2] > 0 and tr[i, 1] > 0: a[3] += (tr[i, 2] - tr[i, 1]) / 1e3  # weight chunks + mma (gemv) / sampling
    if tr[i, 2] > 0: a[4] += (tr[i, 3] - tr[i, 2]) / 1e3              # partial-sum pass + epilogue
print(f"frame total {(tr[-1,3]-tr[0,0])/1e3:.1f} us over {nph} phases (CTA 0 timeline)")
print(f"{'phase':10s} {'n':>4s} {'total us':>9s} {'staged':>9s} {'stream':>9s} {'epilogue':>9s}   (avg per phase)")
for k, a in agg.items():
    print(f"{k:10s} {a[0]:4d} {a[1]/a[0]:9.2f} {a[2]/a[0]:9.2f} {a[3]/a[0]:9.2f} {a[4]/a[0]:9.2f}   total {a[1]:8.1f}")
