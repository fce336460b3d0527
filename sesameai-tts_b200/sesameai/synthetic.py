"""Synthetic weights and prompts for offline parity tests and benchmarks.

No checkpoint, tokenizer or Mimi weight file is reachable offline (the reference
pulls them from the HF hub: ``sesameai/generator.py:28-29,52,338``), so tests and
``bench.py`` use seeded random-init weights of the CSM-1B architecture and
random token ids.

The fill is a counter-based integer hash evaluated with torch integer ops, so the
SAME bits come out on CPU and on CUDA: golden vectors generated in a CPU-only
container stay valid on the GPU box, and the oracle (CPU) and the CUDA path are
always handed identical parameters.

``plant_greedy_structure`` engineers the *decision margin* of every argmax so
that greedy (topk=1) decoding is well-posed in bf16 (SURVEY.md Appendix C.1:
random bf16 logits have an exact top-2 tie in ~2.5-3 % of argmaxes, which the
reference's ``sample_topk`` resolves by RNG, ``sesameai/models.py:81-86``).
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Tuple

import torch

_M32 = 0xFFFFFFFF
_CHUNK = 1 << 24


def _mix(x: torch.Tensor) -> torch.Tensor:
    # 32-bit avalanche hash carried in int64 lanes; both multipliers are < 2**31 so
    # the products never leave the signed 64-bit range.
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & _M32
    x = x ^ (x >> 15)
    x = (x * 0x2C1B3C6D) & _M32
    x = x ^ (x >> 16)
    return x


def hash_uniform_(t: torch.Tensor, seed: int, stream: int, scale: float) -> torch.Tensor:
    """Fill ``t`` in place with U(-scale, scale) values that depend only on
    (flat index, seed, stream) -- identical bits on every device."""
    flat = t.view(-1)
    n = flat.numel()
    key = (seed * 0x9E3779B1 + stream * 0x85EBCA6B + 0x165667B1) & _M32
    for lo in range(0, n, _CHUNK):
        hi = min(n, lo + _CHUNK)
        idx = torch.arange(lo, hi, dtype=torch.int64, device=t.device)
        h = _mix((idx + key) & _M32)
        h = _mix(h ^ ((idx >> 32) + key))
        u = (h >> 8).to(torch.float32) * (1.0 / (1 << 24))  # [0,1) on a 2^-24 grid, exact in fp32
        flat[lo:hi] = ((u - 0.5) * (2.0 * scale)).to(t.dtype)
    return t


def hash_ints(n: int, seed: int, stream: int, high: int, device="cpu") -> torch.Tensor:
    """``n`` int64 values in [0, high), device independent."""
    idx = torch.arange(n, dtype=torch.int64, device=device)
    key = (seed * 0x9E3779B1 + stream * 0x85EBCA6B + 0x27D4EB2F) & _M32
    h = _mix(_mix((idx + key) & _M32) ^ key)
    return h % high


def hash_permutation(n: int, seed: int, stream: int) -> torch.Tensor:
    """A fixed permutation of 0..n-1 (argsort of hashed keys; ties broken by index)."""
    keys = hash_ints(n, seed, stream, 1 << 31) * n + torch.arange(n, dtype=torch.int64)
    return torch.argsort(keys)


def _stream_id(name: str) -> int:
    h = 2166136261
    for ch in name.encode():
        h = ((h ^ ch) * 16777619) & _M32
    return h


@torch.no_grad()
def init_random_weights(model: torch.nn.Module, seed: int = 1234, residual_out_scale: float = 1.0) -> None:
    """Seeded random init at realistic scales for every parameter of a CSM model
    (works on the product ``Model``, the oracle model and the reference ``Model``:
    they share state-dict keys).  Linear weights U(+-1/sqrt(fan_in)), embeddings
    U(+-sqrt(3)) (unit variance), norm scales 1 +- 0.1.  ``audio_head`` -- which the
    reference leaves uninitialised (``sesameai/models.py:118``) -- gets U(+-1/sqrt(1024)).
    ``residual_out_scale`` scales ``output_proj`` and ``mlp.w2`` (1.0 = plain random)."""
    for name, p in sorted(model.state_dict().items()):
        if not torch.is_floating_point(p) or name.endswith("causal_mask"):
            continue
        sid = _stream_id(name)
        if name.endswith(".scale"):
            hash_uniform_(p, seed, sid, 0.1)
            p.add_(1.0)
        elif "embeddings" in name:
            hash_uniform_(p, seed, sid, 3.0 ** 0.5)
        elif name == "audio_head":
            hash_uniform_(p, seed, sid, 1.0 / p.shape[1] ** 0.5)
        elif p.dim() == 2:
            s = 1.0 / p.shape[1] ** 0.5
            if name.endswith("output_proj.weight") or name.endswith("mlp.w2.weight"):
                s *= residual_out_scale
            hash_uniform_(p, seed, sid, s)
        else:  # pragma: no cover - no such parameter in CSM
            hash_uniform_(p, seed, sid, 0.02)


@torch.no_grad()
def plant_greedy_structure(model: torch.nn.Module, seed: int = 1234, logit_peak: float = 12.0,
                           c0_boost: float = 4.0) -> Dict[str, torch.Tensor]:
    """Overwrite the heads (and the text-embedding table) so every greedy argmax has
    a wide bf16 margin, while every kernel still runs on dense full-size data.

    * ``audio_head[i-1][:, j] = a * P @ A_{i-1}[pi_i(j)]``: the depth decoder's input at
      step i is ``P @ A_{i-1}[c_{i-1}]`` (``sesameai/models.py:173,178``), so token
      ``pi_i^-1(c_{i-1})`` wins step i.
    * ``codebook0_head[j] = a * A_31[pi_0(j)]``: the backbone input of an audio frame
      contains ``A_31[c_31]`` (``sesameai/models.py:155-157``); the codebook-31 rows of the
      embedding table are scaled by ``c0_boost`` so that this term dominates the 32-way sum.
    * ``text_embeddings[t] = A_31[t mod V]`` so the first frame after a text-only
      prompt is decided the same way.
    Call after ``init_random_weights(..., residual_out_scale=0.1)``.  The heads are
    computed in fp64 on the CPU so the planted bits do not depend on the device.
    Returns the permutations (``pi[i]`` for codebook i) for analytic checks."""
    sd = model.state_dict()
    V = model.config.audio_vocab_size
    C = model.config.audio_num_codebooks
    A = sd["audio_embeddings.weight"]
    P = sd["projection.weight"]
    dev, dt = A.device, A.dtype
    A.view(C, V, -1)[C - 1].mul_(c0_boost)
    A64 = A.detach().to("cpu", torch.float64).view(C, V, -1)
    P64 = P.detach().to("cpu", torch.float64)
    perms = {i: hash_permutation(V, seed, 0xA000 + i) for i in range(C)}
    d_bb = A64.shape[-1]
    d_dec = P64.shape[0]

    # codebook 0 head: winner logit ~ a*|A_31[c]|^2 / rms(h_in), |A_31|^2 ~ boost^2 d_bb,
    # rms(h_in) ~ sqrt(C - 1 + boost^2)
    a0 = logit_peak * ((C - 1 + c0_boost ** 2) ** 0.5) / (c0_boost ** 2 * d_bb)
    sd["codebook0_head.weight"].copy_((a0 * A64[C - 1][perms[0]]).to(dt).to(dev))
    # depth heads: winner logit ~ a*|P A|^2 / rms(P A) = a*sqrt(d_dec)*|P A|, |P A| ~ sqrt(d_dec)*rms
    head = sd["audio_head"]
    for i in range(1, C):
        PA = A64[i - 1] @ P64.t()  # [V, d_dec]
        rms = PA.pow(2).mean().sqrt()
        a = logit_peak / (d_dec * rms)
        head[i - 1].copy_((a * PA[perms[i]]).t().to(dt).to(dev))
    T = sd["text_embeddings.weight"]
    rows = torch.arange(T.shape[0], device=dev) % V
    T.copy_(A.view(C, V, -1)[C - 1][rows])
    return perms


def planted_next_frame(perms: Dict[int, torch.Tensor], c31_prev: int) -> torch.Tensor:
    """Analytic greedy frame under ``plant_greedy_structure``: c0 = pi_0^-1(c31_prev),
    c_i = pi_i^-1(c_{i-1})."""
    inv = {i: torch.argsort(p) for i, p in perms.items()}
    out = []
    prev = c31_prev
    for i in range(len(perms)):
        prev = int(inv[i][prev])
        out.append(prev)
    return torch.tensor(out, dtype=torch.int32)


def text_prompt(batch: int, frames: int, seed: int = 4321, text_vocab: int = 128_256, n_cols: int = 33,
                device="cpu") -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Text-only prompt frames as ``Generator._tokenize_text_segment`` lays them out
    (reference ``sesameai/generator.py:63-76``): last column = token id, mask last column."""
    tok = torch.zeros(batch, frames, n_cols, dtype=torch.int64)
    msk = torch.zeros(batch, frames, n_cols, dtype=torch.bool)
    tok[:, :, -1] = hash_ints(batch * frames, seed, 1, text_vocab).view(batch, frames)
    msk[:, :, -1] = True
    pos = torch.arange(frames, dtype=torch.int64).unsqueeze(0).repeat(batch, 1)
    return tok.to(device), msk.to(device), pos.to(device)


def voice_prompt(batch: int, segments: int, text_frames: int, audio_frames: int, tail_text_frames: int,
                 seed: int = 4321, text_vocab: int = 128_256, audio_vocab: int = 2051, n_cols: int = 33,
                 device="cpu") -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Context prompt: ``segments`` x (text frames + audio frames whose last frame is the
    all-zero EOS frame, ``sesameai/generator.py:88-89``) followed by the text to speak."""
    toks, msks = [], []
    for s in range(segments):
        t = torch.zeros(batch, text_frames, n_cols, dtype=torch.int64)
        m = torch.zeros(batch, text_frames, n_cols, dtype=torch.bool)
        t[:, :, -1] = hash_ints(batch * text_frames, seed, 10 + 2 * s, text_vocab).view(batch, text_frames)
        m[:, :, -1] = True
        a = torch.zeros(batch, audio_frames, n_cols, dtype=torch.int64)
        am = torch.zeros(batch, audio_frames, n_cols, dtype=torch.bool)
        a[:, :, :-1] = hash_ints(batch * audio_frames * (n_cols - 1), seed, 11 + 2 * s, audio_vocab).view(
            batch, audio_frames, n_cols - 1)
        a[:, -1, :] = 0
        am[:, :, :-1] = True
        toks += [t, a]
        msks += [m, am]
    t, m, _ = text_prompt(batch, tail_text_frames, seed + 7, text_vocab, n_cols)
    toks.append(t)
    msks.append(m)
    tok = torch.cat(toks, dim=1)
    msk = torch.cat(msks, dim=1)
    pos = torch.arange(tok.shape[1], dtype=torch.int64).unsqueeze(0).repeat(batch, 1)
    return tok.to(device), msk.to(device), pos.to(device)


def exp_noise(steps: int, batch: int, vocab: int, seed: int = 777, dtype=torch.bfloat16, device="cpu") -> torch.Tensor:
    """Shared Exp(1) race noise ``q`` for ``sample_topk`` (``sesameai/models.py:72-74``),
    one ``[batch, vocab]`` tensor per sampling call, in the tensor dtype the reference
    draws it in.  ``-log(1-u)`` is evaluated in fp64 on the CPU then rounded, so the bits
    are device independent; values are clamped away from 0 like torch's ``exponential_``."""
    u = torch.empty(steps * batch * vocab, dtype=torch.float64)
    idx = torch.arange(u.numel(), dtype=torch.int64)
    key = (seed * 0x9E3779B1 + 0x51ED270B) & _M32
    h = _mix(_mix((idx + key) & _M32) ^ (idx >> 32))
    u = (h >> 8).to(torch.float64) * (1.0 / (1 << 24))
    q = (-torch.log1p(-u)).clamp_min(2.0 ** -24)
    return q.view(steps, batch, vocab).to(dtype).to(device)


def named_tiny_flavors() -> Dict[str, Dict[str, int]]:
    """Small architectures with the same structure as llama-1B / llama-100M
    (``sesameai/models.py:10-39``) for fast CPU tests."""
    return {
        "tiny-bb": dict(num_layers=2, num_heads=4, num_kv_heads=2, embed_dim=256, intermediate_dim=512),
        "tiny-dec": dict(num_layers=2, num_heads=2, num_kv_heads=1, embed_dim=256, intermediate_dim=512),
    }


@torch.no_grad()
def init_mimi_weights(codec: torch.nn.Module, seed: int = 2024) -> None:
    """Seeded random init for a Mimi codec state dict (moshi key names; works on the oracle and on
    the product codec).  Codebooks are non-zero (moshi's defaults give all-zero embeddings), layer
    scales are O(0.3) so the transformer visibly contributes, conv/linear weights U(+-1/sqrt(fan_in))."""
    for name, p in sorted(codec.state_dict().items()):
        sid = _stream_id(name)
        if name.endswith("_initialized"):
            p.fill_(1.0)
        elif name.endswith("cluster_usage"):
            hash_uniform_(p, seed, sid, 0.5)
            p.add_(1.5)
        elif name.endswith("embedding_sum"):
            hash_uniform_(p, seed, sid, 1.5)
        elif name.endswith("layer_scale_1.scale") or name.endswith("layer_scale_2.scale"):
            hash_uniform_(p, seed, sid, 0.3)
        elif ".norm1." in name or ".norm2." in name:
            hash_uniform_(p, seed, sid, 0.1)
            if name.endswith("weight"):
                p.add_(1.0)
        elif name.endswith("bias"):
            hash_uniform_(p, seed, sid, 0.05)
        elif p.dim() >= 2:
            if "convtr" in name and "upsample" not in name:  # ConvTranspose1d [in, out, k]: fan_in = in * k / stride(=k/2)
                fan_in = p.shape[0] * 2
            else:
                fan_in = p[0].numel()
            hash_uniform_(p, seed, sid, (3.0 / fan_in) ** 0.5)
        else:  # pragma: no cover
            hash_uniform_(p, seed, sid, 0.02)
