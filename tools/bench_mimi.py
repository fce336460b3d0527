"""Mimi decode timing (BASELINE config 4 shape, B from argv) on the current decode path (MIMI_DECODE=mma: mma.sync kernels)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sesameai-tts_b200")); sys.path.insert(0, ROOT)
import torch
from sesameai import synthetic as syn
from sesameai.mimi import MimiCodec
dev = torch.device("cuda", 0)
codec = MimiCodec(max_frames=760)
syn.init_mimi_weights(codec, 2024)
codec.to(dev)
for B, T in [(1, 125), (int(sys.argv[1]) if len(sys.argv) > 1 else 8, 750)]:
    codes = syn.hash_ints(B * 32 * T, 5, T, 2048, device=dev).view(B, 32, T)
    codec.decode(codes); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); codec.decode(codes); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print(f"mimi decode [{os.environ.get('MIMI_DECODE', 'tcgen05')}] B={B} T={T}: {ms:.2f} ms = {ms / B:.2f} ms/utterance, "
          f"{0.4394e9 * B * T / ms / 1e9:.1f} TFLOP/s, {B * T * 0.08 / (ms / 1e3):.0f} x real time")
