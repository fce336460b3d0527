"""A few batched decode steps (B from PF_B, context 1568) for an ncu launch list (direct path: no graph)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sesameai-tts_b200")); sys.path.insert(0, ROOT)
import torch
import bench
from sesameai import _native, synthetic as syn
dev = torch.device("cuda", 0)
B = int(os.environ.get("PF_B", "32"))
model = bench.build_product(dev, B)
SHORT = os.environ.get("PF_SHORT", "") == "1"
tok, msk, pos = (syn.text_prompt(B, 32, 7, device=dev) if SHORT else syn.voice_prompt(B, 4, 64, 320, 32, seed=3, device=dev))
S = tok.shape[1]
model.reset_caches()
s = model.generate_frame(tok, msk, pos, 0.9, 50)
t = torch.zeros(B, 1, 33, dtype=torch.int64, device=dev)
m = torch.ones(B, 1, 33, dtype=torch.bool, device=dev); m[..., -1] = False
p = torch.full((B, 1), S - 1, dtype=torch.int64, device=dev)
t[:, 0, :32] = s; p.add_(1)
s = model.generate_frame(t, m, p, 0.9, 50, no_graph=True)
torch.cuda.synchronize()
torch.cuda.profiler.start()  # ncu --profile-from-start off: exactly one decode step
t[:, 0, :32] = s; p.add_(1)
s = model.generate_frame(t, m, p, 0.9, 50, no_graph=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
