"""ORACLE / TEST INFRASTRUCTURE ONLY.

Restatement of the torchtune 0.4.0 component builder
``torchtune.models.llama3_2.llama3_2`` reached from the reference at
``sesameai/models.py:11-23,27-39``: one RoPE instance
shared by all layers, bias-free projections, SwiGLU MLP, RMSNorm.
"""
from torch import nn

from ..modules.transformer import (
    FeedForward,
    Llama3ScaledRoPE,
    MultiHeadAttention,
    RMSNorm,
    TransformerDecoder,
    TransformerSelfAttentionLayer,
)


def llama3_2(
    vocab_size: int,
    num_layers: int,
    num_heads: int,
    num_kv_heads: int,
    embed_dim: int,
    max_seq_len: int,
    attn_dropout: float = 0.0,
    rope_base: int = 500_000,
    intermediate_dim=None,
    norm_eps: float = 1e-5,
    scale_factor: int = 32,
) -> TransformerDecoder:
    head_dim = embed_dim // num_heads
    rope = Llama3ScaledRoPE(dim=head_dim, max_seq_len=max_seq_len, base=rope_base, scale_factor=scale_factor)
    layers = []
    for _ in range(num_layers):
        attn = MultiHeadAttention(
            embed_dim=embed_dim,
            num_heads=num_heads,
            num_kv_heads=num_kv_heads,
            head_dim=head_dim,
            q_proj=nn.Linear(embed_dim, num_heads * head_dim, bias=False),
            k_proj=nn.Linear(embed_dim, num_kv_heads * head_dim, bias=False),
            v_proj=nn.Linear(embed_dim, num_kv_heads * head_dim, bias=False),
            output_proj=nn.Linear(embed_dim, embed_dim, bias=False),
            pos_embeddings=rope,
            max_seq_len=max_seq_len,
            attn_dropout=attn_dropout,
        )
        mlp = FeedForward(
            gate_proj=nn.Linear(embed_dim, intermediate_dim, bias=False),
            down_proj=nn.Linear(intermediate_dim, embed_dim, bias=False),
            up_proj=nn.Linear(embed_dim, intermediate_dim, bias=False),
        )
        layers.append(
            TransformerSelfAttentionLayer(
                attn, mlp, sa_norm=RMSNorm(embed_dim, eps=norm_eps), mlp_norm=RMSNorm(embed_dim, eps=norm_eps)
            )
        )
    return TransformerDecoder(
        # The reference replaces both with nn.Identity straight away
        # (sesameai/models.py:48-52) after reading ``embedding_dim``; allocate them on
        # the meta device so the oracle does not spend 2 GB on tensors nobody reads.
        tok_embeddings=nn.Embedding(vocab_size, embed_dim, device="meta"),
        layers=layers,
        max_seq_len=max_seq_len,
        num_heads=num_heads,
        head_dim=head_dim,
        norm=RMSNorm(embed_dim, eps=norm_eps),
        output=nn.Linear(embed_dim, vocab_size, bias=False, device="meta"),
    )
