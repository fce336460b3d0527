"""ORACLE / TEST INFRASTRUCTURE ONLY.

Plain-PyTorch restatement of the torchtune 0.4.0 modules on the CSM hot path
(SURVEY.md Appendix A.1-A.5).  Each class states which published torchtune
class it follows and which reference call site reaches it.  Parameter names
match torchtune's so that reference state-dict keys line up
(``layers.{i}.attn.q_proj.weight`` ... ``norm.scale``).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn


class RMSNorm(nn.Module):
    """torchtune.modules.RMSNorm (0.4.0): fp32 normalise, cast back to the input
    dtype, THEN multiply by ``scale`` (two roundings in bf16)."""

    def __init__(self, dim: int, eps: float = 1e-6) -> None:
        super().__init__()
        self.eps = eps
        self.scale = nn.Parameter(torch.ones(dim))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        xf = x.float()
        normed = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + self.eps)).type_as(x)
        return normed * self.scale


class Llama3ScaledRoPE(nn.Module):
    """torchtune.models.llama3_1.Llama3ScaledRoPE (0.4.0).  cos/sin table is a
    non-persistent buffer, so ``model.to(dtype=bf16)`` (reference
    ``generator.py:343``) rounds it to bf16; the rotation itself is fp32 on
    interleaved pairs."""

    def __init__(
        self,
        dim: int,
        max_seq_len: int = 4096,
        base: int = 10_000,
        scale_factor: int = 8,
        low_freq_factor: int = 1,
        high_freq_factor: int = 4,
        old_context_len: int = 8192,
    ) -> None:
        super().__init__()
        self.dim = dim
        self.base = base
        self.max_seq_len = max_seq_len
        self.scale_factor = scale_factor
        self.low_freq_factor = low_freq_factor
        self.high_freq_factor = high_freq_factor
        self.old_context_len = old_context_len
        self.is_cache_built = False
        self.rope_init()

    def rope_init(self) -> None:
        half = self.dim // 2
        freqs = 1.0 / (self.base ** (torch.arange(0, self.dim, 2)[:half].float() / self.dim))
        theta = self._scale(freqs)
        self.register_buffer("theta", theta, persistent=False)
        pos = torch.arange(self.max_seq_len, dtype=theta.dtype, device=theta.device)
        ang = torch.einsum("i, j -> ij", pos, theta).float()
        self.register_buffer("cache", torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1), persistent=False)
        self.is_cache_built = True

    def _scale(self, freqs: torch.Tensor) -> torch.Tensor:
        lo_wavelen = self.old_context_len / self.low_freq_factor
        hi_wavelen = self.old_context_len / self.high_freq_factor
        out = []
        for f in freqs:
            wavelen = 2 * math.pi / f
            if wavelen < hi_wavelen:
                out.append(f)
            elif wavelen > lo_wavelen:
                out.append(f / self.scale_factor)
            else:
                smooth = (self.old_context_len / wavelen - self.low_freq_factor) / (
                    self.high_freq_factor - self.low_freq_factor
                )
                out.append((1 - smooth) * f / self.scale_factor + smooth * f)
        return torch.tensor(out, dtype=freqs.dtype, device=freqs.device)

    def forward(self, x: torch.Tensor, *, input_pos: Optional[torch.Tensor] = None) -> torch.Tensor:
        # x: [b, s, n_h, h_d]
        s = x.size(1)
        table = self.cache[:s] if input_pos is None else self.cache[input_pos]
        xs = x.float().reshape(*x.shape[:-1], -1, 2)
        table = table.view(-1, xs.size(1), 1, xs.size(3), 2)
        rot = torch.stack(
            [
                xs[..., 0] * table[..., 0] - xs[..., 1] * table[..., 1],
                xs[..., 1] * table[..., 0] + xs[..., 0] * table[..., 1],
            ],
            -1,
        )
        return rot.flatten(3).type_as(x)


class KVCache(nn.Module):
    """torchtune.modules.KVCache (0.4.0): [B, H(expanded), max_seq, hd] buffers,
    write position from an internal ``cache_pos`` counter (not ``input_pos``)."""

    def __init__(self, batch_size: int, max_seq_len: int, num_heads: int, head_dim: int, dtype: torch.dtype) -> None:
        super().__init__()
        shape = (batch_size, num_heads, max_seq_len, head_dim)
        self.register_buffer("k_cache", torch.zeros(shape, dtype=dtype), persistent=False)
        self.register_buffer("v_cache", torch.zeros(shape, dtype=dtype), persistent=False)
        self.register_buffer("cache_pos", torch.arange(0, max_seq_len), persistent=False)
        self.batch_size = batch_size

    @property
    def size(self) -> int:
        return int(self.cache_pos[0].item())

    def reset(self) -> None:
        self.k_cache.zero_()
        self.v_cache.zero_()
        self.cache_pos -= self.size

    def update(self, k_val: torch.Tensor, v_val: torch.Tensor):
        bsz, _, seq_len, _ = k_val.shape
        if bsz > self.k_cache.shape[0]:
            raise ValueError(
                f"The current cache has been setup with a batch size of {self.k_cache.shape[0]}"
                f", but found new key tensors with batch size {k_val.shape[0]}!"
            )
        assert (self.cache_pos[0] + seq_len) <= self.k_cache.shape[2]
        self.k_cache[:, :, self.cache_pos[:seq_len]] = k_val
        self.v_cache[:, :, self.cache_pos[:seq_len]] = v_val
        self.cache_pos.add_(seq_len)
        return self.k_cache, self.v_cache


class MultiHeadAttention(nn.Module):
    """torchtune.modules.MultiHeadAttention (0.4.0), self-attention with GQA:
    q head j shares kv head j // (H/KV); k/v are expanded BEFORE the cache."""

    def __init__(
        self,
        *,
        embed_dim: int,
        num_heads: int,
        num_kv_heads: int,
        head_dim: int,
        q_proj: nn.Module,
        k_proj: nn.Module,
        v_proj: nn.Module,
        output_proj: nn.Module,
        pos_embeddings: Optional[nn.Module] = None,
        max_seq_len: int = 4096,
        attn_dropout: float = 0.0,
    ) -> None:
        super().__init__()
        self.embed_dim, self.num_heads, self.num_kv_heads, self.head_dim = embed_dim, num_heads, num_kv_heads, head_dim
        self.max_seq_len = max_seq_len
        self.attn_dropout = attn_dropout
        self.q_proj, self.k_proj, self.v_proj, self.output_proj = q_proj, k_proj, v_proj, output_proj
        self.pos_embeddings = pos_embeddings
        self.kv_cache: Optional[KVCache] = None
        self.cache_enabled = False

    def setup_cache(self, batch_size: int, dtype: torch.dtype, max_seq_len: int) -> None:
        self.kv_cache = KVCache(batch_size, max_seq_len, self.num_heads, self.head_dim, dtype)
        self.cache_enabled = True

    def reset_cache(self) -> None:
        if self.kv_cache is None:
            raise RuntimeError("Key value caches are not setup. Call ``setup_caches()`` first.")
        self.kv_cache.reset()

    def forward(self, x, y=None, *, mask=None, input_pos=None):
        b, s, _ = x.shape
        rep = self.num_heads // self.num_kv_heads
        q = self.q_proj(x).view(b, s, self.num_heads, self.head_dim)
        if self.pos_embeddings is not None:
            q = self.pos_embeddings(q, input_pos=input_pos)
        q = q.transpose(1, 2)

        k = self.k_proj(y).view(b, s, -1, self.head_dim)
        v = self.v_proj(y)
        if self.pos_embeddings is not None:
            k = self.pos_embeddings(k, input_pos=input_pos)
        k = k.view(b, s, self.num_kv_heads, 1, self.head_dim)
        v = v.view(b, s, self.num_kv_heads, 1, self.head_dim)
        if rep != 1:
            k = k.expand(b, s, self.num_kv_heads, rep, self.head_dim)
            v = v.expand(b, s, self.num_kv_heads, rep, self.head_dim)
        k = k.reshape(b, s, -1, self.head_dim).transpose(1, 2)
        v = v.reshape(b, s, -1, self.head_dim).transpose(1, 2)
        if self.kv_cache is not None and self.cache_enabled:
            k, v = self.kv_cache.update(k, v)

        attn_mask = mask[:, None, :, :] if mask is not None else None
        out = F.scaled_dot_product_attention(
            q, k, v, attn_mask=attn_mask, dropout_p=0.0, is_causal=self.kv_cache is None and mask is None
        )
        return self.output_proj(out.transpose(1, 2).contiguous().view(b, s, -1))


class FeedForward(nn.Module):
    """torchtune.modules.FeedForward: w2(silu(w1 x) * w3 x); w1=gate, w3=up, w2=down."""

    def __init__(self, *, gate_proj: nn.Module, down_proj: nn.Module, up_proj: Optional[nn.Module] = None) -> None:
        super().__init__()
        self.w1, self.w2, self.w3 = gate_proj, down_proj, up_proj
        self.activation = nn.SiLU()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.w2(self.activation(self.w1(x)) * self.w3(x))


class TransformerSelfAttentionLayer(nn.Module):
    """torchtune.modules.TransformerSelfAttentionLayer (pre-norm, Identity scales)."""

    def __init__(self, attn: MultiHeadAttention, mlp: nn.Module, *, sa_norm: nn.Module, mlp_norm: nn.Module) -> None:
        super().__init__()
        self.attn, self.mlp, self.sa_norm, self.mlp_norm = attn, mlp, sa_norm, mlp_norm
        self.sa_scale, self.mlp_scale = nn.Identity(), nn.Identity()

    def setup_cache(self, batch_size, dtype, *, encoder_max_seq_len=None, decoder_max_seq_len=None) -> None:
        self.attn.setup_cache(batch_size, dtype, max_seq_len=decoder_max_seq_len)

    def reset_cache(self) -> None:
        self.attn.reset_cache()

    def forward(self, x, *, mask=None, input_pos=None, **kwargs):
        n = self.sa_norm(x)
        h = self.sa_scale(self.attn(n, n, mask=mask, input_pos=input_pos)) + x
        return h + self.mlp_scale(self.mlp(self.mlp_norm(h)))


class TransformerDecoder(nn.Module):
    """torchtune.modules.TransformerDecoder (0.4.0): returns ``output(norm(h)).float()``."""

    def __init__(self, *, tok_embeddings, layers, max_seq_len, num_heads, head_dim, norm, output) -> None:
        super().__init__()
        self.tok_embeddings = tok_embeddings
        self.layers = layers if isinstance(layers, nn.ModuleList) else nn.ModuleList(layers)
        self.norm, self.output = norm, output
        self.max_seq_len, self.num_heads, self.head_dim = max_seq_len, num_heads, head_dim
        self.decoder_max_cache_seq_len: Optional[int] = None

    def setup_caches(self, batch_size, dtype, *, encoder_max_seq_len=None, decoder_max_seq_len=None) -> None:
        self.decoder_max_cache_seq_len = decoder_max_seq_len if decoder_max_seq_len is not None else self.max_seq_len
        for layer in self.layers:
            layer.setup_cache(batch_size, dtype, decoder_max_seq_len=self.decoder_max_cache_seq_len)

    def caches_are_enabled(self) -> bool:
        return self.layers[0].attn.kv_cache is not None

    def reset_caches(self) -> None:
        if not self.caches_are_enabled():
            raise RuntimeError("Key value caches are not setup. Call ``setup_caches()`` first.")
        for layer in self.layers:
            layer.reset_cache()

    def forward(self, tokens, *, mask=None, encoder_input=None, encoder_mask=None, input_pos=None):
        seq_len = tokens.shape[1]
        if seq_len > self.max_seq_len:
            raise ValueError(f"seq_len ({seq_len}) of input tensor should be smaller than max_seq_len ({self.max_seq_len})")
        if self.caches_are_enabled():
            if mask is None:
                raise ValueError("KV-caches for self-attention layers are setup for inference mode, causal masks must be provided!")
            if input_pos is None:
                raise ValueError("Caches are setup, but the position of input token is missing")
        h = self.tok_embeddings(tokens)
        for layer in self.layers:
            h = layer(h, mask=mask, input_pos=input_pos)
        return self.output(self.norm(h)).float()
