"""Secondary measurements (not the headline bench line): prompt prefill on the tensor cores,
batched decode through the per-op graph path, Mimi decode.  Writes one JSON object."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sesameai-tts_b200"))
sys.path.insert(0, ROOT)
import torch

import bench
from sesameai import _native, synthetic as syn
from sesameai.mimi import MimiCodec

dev = torch.device("cuda", 0)
out = {}
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}


def ev_time(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


# ---- prefill: S-frame voice prompt, tensor-core path vs small-row path ---------------------------
ONLY = os.environ.get("BENCH_EXTRA_ONLY", "")  # "mimi": skip the language-model part
for B in (() if ONLY == "mimi" else (1, 8, 32)):
    model = bench.build_product(dev, B)
    tok, msk, pos = syn.voice_prompt(B, 4, 64, 320, 32, seed=3, device=dev)  # 1568 frames (BASELINE config 3 prompt)
    S = tok.shape[1]
    def run(mode):
        model.reset_caches()
        model.generate_frame(tok, msk, pos, 0.9, 50, prefill=mode)
    t_tc = ev_time(lambda: run(_native.PREFILL_TENSOR))
    flop = B * (2 * 973_146_112 * S + 2 * S * S * 2048 * 16 / 2)
    out[f"prefill_B{B}_S{S}"] = {"ms_tensor_core": t_tc, "tokens_per_s": B * S / t_tc * 1e3,
                                  "tflops": flop / t_tc / 1e9, "frac_of_bf16_peak": flop / t_tc / 1e9 / peaks.get("bf16_tflops_sustained", 1384.0)}
    if B == 1 and "--small-row" in sys.argv:
        out[f"prefill_B{B}_S{S}"]["ms_small_row"] = ev_time(lambda: run(_native.PREFILL_SMALL_ROW), reps=1, warm=0)
    # ---- batched decode (graph path for B > 1) ------------------------------------------------------
    t = torch.zeros(B, 1, 33, dtype=torch.int64, device=dev)
    m = torch.ones(B, 1, 33, dtype=torch.bool, device=dev); m[..., -1] = False
    p = torch.full((B, 1), S - 1, dtype=torch.int64, device=dev)
    model.reset_caches()
    s = model.generate_frame(tok, msk, pos, 0.9, 50)
    def step():
        global s
        t[:, 0, :32] = s; p.add_(1)
        s = model.generate_frame(t, m, p, 0.9, 50)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        step()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    out[f"decode_B{B}_ctx{S}"] = {"ms_per_step": ms, "frames_per_s": B / ms * 1e3,
                                   "hbm_frac": bench.bytes_per_frame(B, S + 12) / (ms / 1e3) / 1e9 / peaks.get("hbm_gbs", 6553.0)}
    del model
    torch.cuda.empty_cache()

# ---- Mimi decode ------------------------------------------------------------------------------------------
codec = MimiCodec(max_frames=760)
syn.init_mimi_weights(codec, 2024)
codec.to(dev)
for B, T in ((1, 125), (4, 750)):
    codes = syn.hash_ints(B * 32 * T, 5, T, 2048, device=dev).view(B, 32, T)
    ms = ev_time(lambda: codec.decode(codes), reps=2)
    flop = 0.4394e9 * B * T
    out[f"mimi_decode_B{B}_T{T}"] = {"ms": ms, "audio_s": B * T * 0.08, "x_realtime": B * T * 0.08 / (ms / 1e3),
                                      "tflops": flop / ms / 1e9}
print(json.dumps(out, indent=1))
