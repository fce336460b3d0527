"""ORACLE / TEST INFRASTRUCTURE ONLY -- the product path never imports this file.

CPU/PyTorch restatement of the CSM-1B frame-generation algorithm of the
reference (``/root/reference/sesameai/models.py``), written so that it executes
the same sequence of torch ops with the same dtypes and therefore reproduces the
reference bit-for-bit on the CPU.  The transformer blocks come from
``oracle/shim/torchtune`` (restated torchtune 0.4.0; not vendored by the
reference, ``requirements.txt:7``).

Pinning (see DESIGN.md "Oracle"):
  * ``tests/test_oracle_vs_reference.py`` (runs where ``/root/reference`` exists)
    imports the UNMODIFIED reference ``sesameai/models.py`` on top of the shim and
    requires token- and logit-identical output from this restatement;
  * ``tests/golden/*.pt`` hold outputs of that reference run
    (``tests/golden/make_golden.py`` is the generating script);
  * ``tests/test_oracle_pin_hf.py`` checks the fp32 mathematics against the
    independent ``transformers`` CSM port.
The reference itself has no tests or golden vectors, so the
bf16 rounding points of torchtune 0.4.0 are restated from its published source.
"""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional

import torch
from torch import nn

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")
if _SHIM not in sys.path:
    sys.path.insert(0, _SHIM)

from torchtune.models import llama3_2 as _tt_llama  # noqa: E402  (the shim)


@dataclass
class OracleArgs:
    """Mirrors ``ModelArgs`` (``sesameai/models.py:90-96``)."""

    backbone_flavor: str = "llama-1B"
    decoder_flavor: str = "llama-100M"
    text_vocab_size: int = 128_256
    audio_vocab_size: int = 2051
    audio_num_codebooks: int = 32


# ``sesameai/models.py:10-39``: the two llama3_2 flavours.  Extra (tiny) flavours can be
# registered by tests; they keep every other builder argument.
ARCH: Dict[str, Dict[str, int]] = {
    "llama-1B": dict(num_layers=16, num_heads=32, num_kv_heads=8, embed_dim=2048, intermediate_dim=8192),
    "llama-100M": dict(num_layers=4, num_heads=8, num_kv_heads=2, embed_dim=1024, intermediate_dim=8192),
}


def build_stack(flavor: str):
    a = ARCH[flavor]
    net = _tt_llama.llama3_2(
        vocab_size=128_256, max_seq_len=2048, attn_dropout=0.0, norm_eps=1e-5, rope_base=500_000, scale_factor=32, **a
    )
    # ``_prepare_transformer`` (sesameai/models.py:48-52)
    net.tok_embeddings = nn.Identity()
    net.output = nn.Identity()
    return net, a["embed_dim"]


def exp_race_argmax(probs: torch.Tensor, q: Optional[torch.Tensor]) -> torch.Tensor:
    """``_multinomial_sample_one_no_sync`` (sesameai/models.py:72-74); ``q`` may be
    supplied so two implementations share the Exp(1) noise."""
    if q is None:
        q = torch.empty_like(probs).exponential_(1)
    return torch.argmax(probs / q, dim=-1, keepdim=True).to(dtype=torch.int)


def oracle_sample_topk(logits: torch.Tensor, topk: int, temperature: float, q: Optional[torch.Tensor] = None,
                       autocast_cuda: bool = False):
    """``sample_topk`` (sesameai/models.py:77-87): ties with the k-th value are kept.  ``autocast_cuda``: the
    op dtypes of ``torch.autocast("cuda", bfloat16)`` (tts_service.py:192-194): log_softmax / softmax are on
    autocast's fp32 list, so the probabilities and the Exp(1) race are fp32 (SURVEY.md 8c mode ii)."""
    x = logits / temperature
    kth = torch.topk(x, topk)[0][..., -1, None]
    x = x.masked_fill(x < kth, -float("Inf"))
    if autocast_cuda:
        x = torch.nn.functional.log_softmax(x, dim=-1, dtype=torch.float32)
        p = torch.nn.functional.softmax(x, dim=-1, dtype=torch.float32)
        return exp_race_argmax(p, None if q is None else q.float())
    x = torch.nn.functional.log_softmax(x, dim=-1)
    p = torch.nn.functional.softmax(x, dim=-1)
    return exp_race_argmax(p, q)


class OracleCSM(nn.Module):
    """State-dict compatible with the reference ``Model`` (sesameai/models.py:99-118)."""

    def __init__(self, config: OracleArgs):
        super().__init__()
        self.config = config
        self.backbone, d_bb = build_stack(config.backbone_flavor)
        self.decoder, d_dec = build_stack(config.decoder_flavor)
        V, C = config.audio_vocab_size, config.audio_num_codebooks
        self.text_embeddings = nn.Embedding(config.text_vocab_size, d_bb)
        self.audio_embeddings = nn.Embedding(V * C, d_bb)
        self.projection = nn.Linear(d_bb, d_dec, bias=False)
        self.codebook0_head = nn.Linear(d_bb, V, bias=False)
        self.audio_head = nn.Parameter(torch.empty(C - 1, d_dec, V))
        # Mode (ii) of SURVEY.md 8c: the op dtypes of ``torch.autocast("cuda", dtype=bfloat16)`` around the frame
        # loop (tts_service.py:192-194), emulated on the CPU (CPU autocast has different op lists).  What changes
        # against the plain mode: ``sum`` is on autocast's fp32 list, so the embedding sum -- and with it the whole
        # BACKBONE residual stream and its RMSNorm outputs -- is fp32; every linear layer (lower-precision list)
        # rounds its input to bf16 once; log_softmax / softmax and the Exp(1) race are fp32.  The depth decoder
        # starts from bf16 inputs and stays bf16.  The forward pre-hooks below are the "cast to bf16" of autocast's
        # linear; they are no-ops in the plain mode, where every input is bf16 already.
        self.autocast_cuda = False
        for mod in self.modules():
            if isinstance(mod, nn.Linear):
                mod.register_forward_pre_hook(lambda m, args: (args[0].to(m.weight.dtype),) + tuple(args[1:]))

    # -- sesameai/models.py:120-130
    def setup_caches(self, max_batch_size: int) -> None:
        p = next(self.parameters())
        with p.device:
            self.backbone.setup_caches(max_batch_size, p.dtype)
            self.decoder.setup_caches(max_batch_size, p.dtype, decoder_max_seq_len=self.config.audio_num_codebooks)
        tri = lambda n: torch.tril(torch.ones(n, n, dtype=torch.bool, device=p.device))  # noqa: E731
        self.register_buffer("backbone_causal_mask", tri(self.backbone.max_seq_len))
        self.register_buffer("decoder_causal_mask", tri(self.config.audio_num_codebooks))

    # -- sesameai/models.py:186-188
    def reset_caches(self) -> None:
        self.backbone.reset_caches()
        self.decoder.reset_caches()

    # -- sesameai/models.py:190-203 + :155-157
    def embed_frame_inputs(self, tokens: torch.Tensor, tokens_mask: torch.Tensor) -> torch.Tensor:
        V, C = self.config.audio_vocab_size, self.config.audio_num_codebooks
        txt = self.text_embeddings(tokens[:, :, -1]).unsqueeze(-2)
        aud_idx = tokens[:, :, :-1] + V * torch.arange(C, device=tokens.device)
        aud = self.audio_embeddings(aud_idx.view(-1)).reshape(tokens.size(0), tokens.size(1), C, -1)
        stacked = torch.cat([aud, txt], dim=-2)
        masked = stacked * tokens_mask.unsqueeze(-1)
        return masked.sum(dim=2, dtype=torch.float32) if self.autocast_cuda else masked.sum(dim=2)

    def embed_audio(self, codebook: int, tok: torch.Tensor) -> torch.Tensor:
        return self.audio_embeddings(tok + codebook * self.config.audio_vocab_size)

    # -- sesameai/models.py:132-184
    def generate_frame(
        self,
        tokens: torch.Tensor,
        tokens_mask: torch.Tensor,
        input_pos: torch.Tensor,
        temperature: float,
        topk: int,
        *,
        noise: Optional[torch.Tensor] = None,  # [32, B, V] Exp(1) draws, one per codebook
        forced: Optional[torch.Tensor] = None,  # [B, 32] teacher-forced tokens
        record: Optional[Dict[str, List[torch.Tensor]]] = None,
    ) -> torch.Tensor:
        dt = next(self.parameters()).dtype
        C = self.config.audio_num_codebooks
        assert self.backbone.caches_are_enabled(), "backbone caches are not enabled"
        bb_mask = self.backbone_causal_mask[input_pos, :]
        h = self.embed_frame_inputs(tokens, tokens_mask)
        h = self.backbone(h, input_pos=input_pos, mask=bb_mask).to(dtype=dt)
        last_h = h[:, -1, :]

        def pick(logits: torch.Tensor, i: int) -> torch.Tensor:
            s = oracle_sample_topk(logits, topk, temperature, None if noise is None else noise[i], self.autocast_cuda)
            if record is not None:
                record.setdefault("logits", []).append(logits.detach().clone())
                record.setdefault("sampled", []).append(s.detach().clone())
            return s if forced is None else forced[:, i : i + 1].to(torch.int)

        c = pick(self.codebook0_head(last_h), 0)
        frame = c.clone()
        cur = torch.cat([last_h.unsqueeze(1), self.embed_audio(0, c)], dim=1)
        pos = torch.arange(0, cur.size(1), device=cur.device).unsqueeze(0).repeat(cur.size(0), 1)
        self.decoder.reset_caches()
        for i in range(1, C):
            d_mask = self.decoder_causal_mask[pos, :]
            dh = self.decoder(self.projection(cur), input_pos=pos, mask=d_mask).to(dtype=dt)
            c = pick(torch.mm(dh[:, -1, :], self.audio_head[i - 1]), i)
            cur = self.embed_audio(i, c)
            frame = torch.cat([frame, c], dim=1)
            pos = pos[:, -1:] + 1
        return frame


@torch.inference_mode()
def oracle_frame_loop(
    model,
    tokens: torch.Tensor,
    tokens_mask: torch.Tensor,
    input_pos: torch.Tensor,
    max_frames: int,
    temperature: float,
    topk: int,
    *,
    stop_on_eos: bool = True,
    frame_fn: Optional[Callable] = None,
) -> List[torch.Tensor]:
    """The generation loop of ``Generator.generate`` (sesameai/generator.py:255,283-294;
    twin at tts_service.py:224-241), batch-generalised the way the reference's own tensor
    ops generalise: EOS = every code of every stream is 0.  ``model`` is anything with
    ``reset_caches``/``generate_frame`` (oracle, reference or product)."""
    model.reset_caches()
    out: List[torch.Tensor] = []
    cur_t, cur_m, cur_p = tokens, tokens_mask, input_pos
    B = tokens.size(0)
    dev = tokens.device
    for i in range(max_frames):
        if frame_fn is not None:
            s = frame_fn(i, cur_t, cur_m, cur_p)
        else:
            s = model.generate_frame(cur_t, cur_m, cur_p, temperature, topk)
        if stop_on_eos and torch.all(s == 0):
            break
        out.append(s)
        cur_t = torch.cat([s, torch.zeros(B, 1).long().to(dev)], dim=1).unsqueeze(1)
        cur_m = torch.cat([torch.ones_like(s).bool(), torch.zeros(B, 1).bool().to(dev)], dim=1).unsqueeze(1)
        cur_p = cur_p[:, -1:] + 1
    return out
