"""Oracle restatement of torchtune.modules (0.4.0). Test infrastructure only."""
from . import transformer  # noqa: F401
from .transformer import (  # noqa: F401
    FeedForward,
    KVCache,
    Llama3ScaledRoPE,
    MultiHeadAttention,
    RMSNorm,
    TransformerDecoder,
    TransformerSelfAttentionLayer,
)
