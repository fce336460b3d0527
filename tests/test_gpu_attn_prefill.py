"""Tensor-core prompt attention (k_attn_flash64) vs a plain fp32 PyTorch reference of the same op:
causal attention of prompt rows over the GQA-compact KV cache (torchtune MultiHeadAttention + the
reference's causal mask rows, sesameai/models.py:55-69)."""
import pytest
import torch

from sesameai import _native
from sesameai import synthetic as syn

pytestmark = pytest.mark.gpu


def _case(B, chunk, cache_len, heads, kv_heads, slots, seed, poison=True):
    q = torch.empty(B * chunk, heads * 64, device="cuda")
    kc = torch.empty(B, kv_heads, slots, 64, device="cuda")
    vc = torch.empty(B, kv_heads, slots, 64, device="cuda")
    syn.hash_uniform_(q, seed, 1, 2.0)
    syn.hash_uniform_(kc, seed, 2, 2.0)
    syn.hash_uniform_(vc, seed, 3, 1.0)
    q, kc, vc = q.to(torch.bfloat16), kc.to(torch.bfloat16), vc.to(torch.bfloat16)
    used = cache_len + chunk
    if poison and used < slots:  # never-written cache rows may hold anything, Inf and NaN included
        kc[:, :, used:] = float("nan")
        vc[:, :, used:] = float("inf")
    slot = (cache_len + torch.arange(chunk, device="cuda", dtype=torch.int32)).repeat(B).contiguous()
    out = torch.empty_like(q)
    _native.check(_native.lib().csm_k_attn_prefill(q.data_ptr(), kc.data_ptr(), vc.data_ptr(), slot.data_ptr(), B, chunk, heads,
                                                   kv_heads, slots, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    # fp32 reference
    grp = heads // kv_heads
    qf = q.float().view(B, chunk, heads, 64).permute(0, 2, 1, 3)                      # [B, H, chunk, 64]
    kf = kc.float()[:, :, :used].repeat_interleave(grp, dim=1)                        # [B, H, used, 64]
    vf = vc.float()[:, :, :used].repeat_interleave(grp, dim=1)
    s = qf @ kf.transpose(-1, -2) * 0.125
    keys = torch.arange(used, device="cuda")[None, :]
    vis = keys <= (cache_len + torch.arange(chunk, device="cuda"))[:, None]
    s = s.masked_fill(~vis, float("-inf"))
    ref = (torch.softmax(s, dim=-1) @ vf).permute(0, 2, 1, 3).reshape(B * chunk, heads * 64)
    return out.float(), ref


@pytest.mark.parametrize("B,chunk,cache_len,heads,kv_heads,slots", [
    (1, 64, 0, 4, 2, 128),        # one full tile
    (2, 300, 0, 8, 2, 512),       # ragged last tile, two streams
    (1, 17, 40, 4, 4, 128),       # short chunk on top of a cache, partial key tile
    (3, 130, 70, 32, 8, 256),     # CSM-1B backbone head layout
    (1, 1567, 0, 32, 8, 2048),    # BASELINE config 3 prompt length
])
def test_prefill_attention_matches_fp32_reference(B, chunk, cache_len, heads, kv_heads, slots):
    got, ref = _case(B, chunk, cache_len, heads, kv_heads, slots, seed=B * 1000 + chunk)
    assert torch.isfinite(got).all()
    # P is rounded to bf16 before the second product and the result to bf16: a few 2^-9 relative steps
    err = (got - ref).abs().max().item()
    assert err <= 3 * 2.0 ** -8 * max(1.0, ref.abs().max().item()), err
    cos = torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
    assert cos >= 0.9999, cos
