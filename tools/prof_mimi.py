"""One 60 s Mimi decode for an ncu launch list (ncu --profile-from-start off)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sesameai-tts_b200")); sys.path.insert(0, ROOT)
import torch
from sesameai import synthetic as syn
from sesameai.mimi import MimiCodec
dev = torch.device("cuda", 0)
codec = MimiCodec(max_frames=760)
syn.init_mimi_weights(codec, 2024)
codec.to(dev)
codes = syn.hash_ints(32 * 750, 5, 750, 2048, device=dev).view(1, 32, 750)
codec.decode(codes); torch.cuda.synchronize()
torch.cuda.profiler.start()
codec.decode(codes); torch.cuda.synchronize()
torch.cuda.profiler.stop()
