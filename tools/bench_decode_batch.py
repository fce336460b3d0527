"""Batched decode step time at context 1568 (graph path), B from argv."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sesameai-tts_b200")); sys.path.insert(0, ROOT)
import torch
import bench
from sesameai import synthetic as syn
dev = torch.device("cuda", 0)
SHORT = os.environ.get("PF_SHORT", "") == "1"  # 32-frame text prompt (config 2 / 5 context) instead of the 1568-frame voice prompt
for B in [int(a) for a in sys.argv[1:]] or [8, 32]:
    model = bench.build_product(dev, B)
    tok, msk, pos = (syn.text_prompt(B, 32, 7, device=dev) if SHORT else syn.voice_prompt(B, 4, 64, 320, 32, seed=3, device=dev))
    S = tok.shape[1]
    model.reset_caches()
    s = model.generate_frame(tok, msk, pos, 0.9, 50)
    t = torch.zeros(B, 1, 33, dtype=torch.int64, device=dev)
    m = torch.ones(B, 1, 33, dtype=torch.bool, device=dev); m[..., -1] = False
    p = torch.full((B, 1), S - 1, dtype=torch.int64, device=dev)
    def step():
        global s
        t[:, 0, :32] = s; p.add_(1)
        s = model.generate_frame(t, m, p, 0.9, 50)
    for _ in range(3): step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): step()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print(f"decode B={B} ctx={S}: {ms:.2f} ms/step = {B / ms * 1e3:.0f} frames/s, "
          f"{bench.bytes_per_frame(B, S + 12) / (ms / 1e3) / 1e9 / 6553.0 * 100:.1f} % of the HBM roofline")
    del model; torch.cuda.empty_cache()
