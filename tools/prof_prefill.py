"""One 1568-frame prompt prefill (B=1) for an ncu launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sesameai-tts_b200")); sys.path.insert(0, ROOT)
import torch
import bench
from sesameai import _native, synthetic as syn
dev = torch.device("cuda", 0)
B = int(os.environ.get("PF_B", "1"))
model = bench.build_product(dev, B)
tok, msk, pos = syn.voice_prompt(B, 4, 64, 320, 32, seed=3, device=dev)
model.reset_caches()
model.generate_frame(tok, msk, pos, 0.9, 50, prefill=_native.PREFILL_TENSOR)
torch.cuda.synchronize()
torch.cuda.profiler.start()  # ncu --profile-from-start off: one prefill + first frame
model.reset_caches()
model.generate_frame(tok, msk, pos, 0.9, 50, prefill=_native.PREFILL_TENSOR)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
