"""Small invocations of the kernels added in the second half of round 2, for `compute-sanitizer --tool memcheck`:
cluster split-K tcgen05 GEMM (DSMEM reduce), 128 x 256 persistent tiles, depth-decoder attention, warp-per-row RMSNorm,
the megakernel's split long-context attention, Mimi's weight-resident tail GEMM with the fused final conv and the batched
transformer.  Sizes are tiny: the sanitizer slows kernels down by one to two orders of magnitude."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sesameai-tts_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
from sesameai import _native, synthetic as syn
from sesameai.mimi import MimiCodec
from helpers import build_product, next_inputs

dev = torch.device("cuda", 0)
# 1. GEMM unit entry: cluster split-K (8 / 4 / 2 slices, three epilogues) and wide persistent tiles
def gemm(N, K, M, epi=0):
    x = torch.empty(N, K, device=dev); w = torch.empty(M, K, device=dev)
    syn.hash_uniform_(x, 1, 1, 1.0); syn.hash_uniform_(w, 1, 2, K ** -0.5)
    x, w = x.to(torch.bfloat16), w.to(torch.bfloat16)
    y = torch.empty(N, M // 2 if epi == 2 else M, dtype=torch.bfloat16, device=dev)
    r = torch.zeros_like(y) if epi == 1 else None
    _native.check(_native.lib().csm_k_gemm_tc(x.data_ptr(), w.data_ptr(), N, K, M, y.data_ptr(), epi, r.data_ptr() if r is not None else None,
                                              torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = x.float() @ w.float().t()
    if epi == 0:
        assert (y.float() - ref).abs().max().item() <= 2.0 ** -7 * max(1.0, ref.abs().max().item())
for shape in [(50, 8192, 1024, 0), (128, 2048, 2048, 1), (130, 1024, 512, 2), (300, 1024, 1536, 0), (700, 256, 3072, 0)]:
    gemm(*shape)
print("gemm ok")
# 2. tiny CSM: 40 streams (tcgen05 decode path: cluster split-K, k_attn_dec, k_rmsnorm_rows fall-backs), 3 frames
spec = dict(model_args=dict(backbone_flavor="tiny-bb", decoder_flavor="tiny-dec", text_vocab_size=1000, audio_vocab_size=2051,
                            audio_num_codebooks=32), weight_seed=1234, planted=False, batch=40)
pm, _ = build_product(spec, batch=40)
tok, msk, pos = syn.text_prompt(40, 5, 11, 1000, device="cuda")
for _ in range(3):
    s = pm.generate_frame(tok, msk, pos, 0.9, 50)
    tok, msk, pos = next_inputs(s, pos)
torch.cuda.synchronize(); pm.check_device_error()
print("batched decode ok")
# 3. megakernel with the split attention: 510-frame prompt, 4 frames (slots 510 .. 513)
pm1, _ = build_product(dict(spec, batch=1), batch=1)
tok, msk, pos = syn.voice_prompt(1, 3, 20, 130, 60, seed=6, text_vocab=1000, device="cuda")
for _ in range(4):
    s = pm1.generate_frame(tok, msk, pos, 0.9, 50)
    tok, msk, pos = next_inputs(s, pos)
torch.cuda.synchronize(); pm1.check_device_error()
print("split attention ok")
# 3b. row-batched decode at long context: 3 streams, 300-frame prompt (k_attn_split64_mma + k_attn_combine64), 2 frames
pm3, _ = build_product(dict(spec, batch=3), batch=3)
tok, msk, pos = syn.voice_prompt(3, 3, 20, 70, 30, seed=5, text_vocab=1000, device="cuda")
for _ in range(3):
    s = pm3.generate_frame(tok, msk, pos, 0.9, 50)
    tok, msk, pos = next_inputs(s, pos)
torch.cuda.synchronize(); pm3.check_device_error()
print("long-context batched decode ok")
# 4. Mimi: batched decode of 3 utterances x 4 frames (resident tail GEMM, fused final conv, batched transformer) + a stream
codec = MimiCodec(max_frames=8); syn.init_mimi_weights(codec, 2024); codec.to(dev)
codes = syn.hash_ints(3 * 32 * 4, 5, 4, 2048, device=dev).view(3, 32, 4)
a = codec.decode(codes)
st = codec.streaming()
b = torch.cat([st.decode(codes[:1, :, :2]), st.decode(codes[:1, :, 2:])], dim=-1)
torch.cuda.synchronize()
assert torch.equal(a[:1], b)
print("mimi ok")
