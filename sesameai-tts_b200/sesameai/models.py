"""Drop-in ``sesameai.models`` for B200: same public surface as the reference module
(``/root/reference/sesameai/models.py``) -- ``ModelArgs``, ``FLAVORS``, ``Model`` with
``setup_caches`` / ``generate_frame`` / ``reset_caches``, ``sample_topk`` -- but ``Model`` is a
parameter container whose hot path is one call into libcsm_b200.so (hand-written sm_100a
kernels, ``include/csm_b200.h``).  There is no PyTorch or CPU execution path in here:
``generate_frame`` raises if the model is not on a CUDA device in bf16.

State-dict keys are those of the reference (``text_embeddings.weight``, ``audio_head``,
``backbone.layers.{i}.attn.q_proj.weight``, ..., ``decoder.norm.scale``) so a ``sesame/csm-1b``
checkpoint loads through ``Model.from_pretrained`` unchanged.
"""
from __future__ import annotations

import ctypes
import math
import weakref
from dataclasses import dataclass
from typing import Dict, Optional

import torch
from torch import nn

try:  # checkpoint loading only; absent hub support must not break the hot path
    from huggingface_hub import PyTorchModelHubMixin
except Exception:  # pragma: no cover
    class PyTorchModelHubMixin:  # type: ignore
        def __init_subclass__(cls, **kw):
            super().__init_subclass__()

from . import _native


@dataclass
class StackSpec:
    """Shape of one llama3_2 stack (reference ``models.py:10-39``)."""

    num_layers: int
    num_heads: int
    num_kv_heads: int
    embed_dim: int
    intermediate_dim: int
    max_seq_len: int = 2048
    norm_eps: float = 1e-5
    rope_base: float = 500_000.0
    scale_factor: float = 32.0

    @property
    def head_dim(self) -> int:
        return self.embed_dim // self.num_heads


def llama3_2_1B() -> "TransformerStack":
    return TransformerStack(StackSpec(16, 32, 8, 2048, 8192))


def llama3_2_100M() -> "TransformerStack":
    return TransformerStack(StackSpec(4, 8, 2, 1024, 8192))


FLAVORS = {"llama-1B": llama3_2_1B, "llama-100M": llama3_2_100M}


def register_flavor(name: str, **dims) -> None:
    """Add an architecture (tests use tiny ones); head_dim must be 64 or 128."""
    FLAVORS[name] = lambda: TransformerStack(StackSpec(**dims))


@dataclass
class ModelArgs:
    backbone_flavor: str
    decoder_flavor: str
    text_vocab_size: int
    audio_vocab_size: int
    audio_num_codebooks: int


# ------------------------------------------------------------------------------------------------
# Parameter containers.  They mirror torchtune's module tree so the parameter names match, but
# own no arithmetic: calling them is an error.
# ------------------------------------------------------------------------------------------------
class _Weight(nn.Module):
    def __init__(self, out_features: int, in_features: int):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))

    def forward(self, *a, **k):
        raise RuntimeError("sesameai(B200): layers are parameter containers; use Model.generate_frame")


class _Scale(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.scale = nn.Parameter(torch.ones(dim))


class _Attention(nn.Module):
    def __init__(self, s: StackSpec):
        super().__init__()
        self.q_proj = _Weight(s.num_heads * s.head_dim, s.embed_dim)
        self.k_proj = _Weight(s.num_kv_heads * s.head_dim, s.embed_dim)
        self.v_proj = _Weight(s.num_kv_heads * s.head_dim, s.embed_dim)
        self.output_proj = _Weight(s.embed_dim, s.embed_dim)


class _SwiGLU(nn.Module):
    def __init__(self, s: StackSpec):
        super().__init__()
        self.w1 = _Weight(s.intermediate_dim, s.embed_dim)  # gate
        self.w2 = _Weight(s.embed_dim, s.intermediate_dim)  # down
        self.w3 = _Weight(s.intermediate_dim, s.embed_dim)  # up


class _Block(nn.Module):
    def __init__(self, s: StackSpec):
        super().__init__()
        self.attn = _Attention(s)
        self.mlp = _SwiGLU(s)
        self.sa_norm = _Scale(s.embed_dim)
        self.mlp_norm = _Scale(s.embed_dim)


class TransformerStack(nn.Module):
    """Stands where torchtune's ``TransformerDecoder`` stands in the reference ``Model``."""

    def __init__(self, spec: StackSpec):
        super().__init__()
        self.spec = spec
        self.layers = nn.ModuleList([_Block(spec) for _ in range(spec.num_layers)])
        self.norm = _Scale(spec.embed_dim)
        self.max_seq_len = spec.max_seq_len
        self.num_heads = spec.num_heads
        self.head_dim = spec.head_dim
        # weak back-reference kept out of nn.Module's attribute machinery (a plain attribute
        # would register the owning Model as a sub-module and create a cycle)
        self.__dict__["_owner_ref"] = None

    def caches_are_enabled(self) -> bool:
        ref = self.__dict__.get("_owner_ref")
        owner = ref() if ref is not None else None
        return owner is not None and owner._ctx is not None

    def forward(self, *a, **k):
        raise RuntimeError("sesameai(B200): the transformer runs inside libcsm_b200; use Model.generate_frame")

    def rope_table(self, n_pos: int) -> torch.Tensor:
        """Llama3ScaledRoPE cache [n_pos, hd/2, 2] = (cos, sin), built in fp32 on the CPU exactly as
        torchtune 0.4.0 does and rounded to bf16 the way ``model.to(dtype=bfloat16)`` rounds the
        reference's buffer (reference ``generator.py:343``)."""
        s = self.spec
        hd = s.head_dim
        freqs = 1.0 / (s.rope_base ** (torch.arange(0, hd, 2)[: hd // 2].float() / hd))
        old_len, lo_f, hi_f = 8192, 1, 4
        scaled = []
        for f in freqs:
            wavelen = 2 * math.pi / f
            if wavelen < old_len / hi_f:
                scaled.append(f)
            elif wavelen > old_len / lo_f:
                scaled.append(f / s.scale_factor)
            else:
                smooth = (old_len / wavelen - lo_f) / (hi_f - lo_f)
                scaled.append((1 - smooth) * f / s.scale_factor + smooth * f)
        theta = torch.tensor(scaled, dtype=freqs.dtype)
        ang = torch.einsum("i, j -> ij", torch.arange(n_pos, dtype=theta.dtype), theta).float()
        return torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).to(torch.bfloat16)


def _multinomial_sample_one_no_sync(probs):
    q = torch.empty_like(probs).exponential_(1)
    return torch.argmax(probs / q, dim=-1, keepdim=True).to(dtype=torch.int)


def sample_topk(logits: torch.Tensor, topk: int, temperature: float) -> torch.Tensor:
    """Same contract as the reference helper (``models.py:77-87``) for callers that import it;
    on CUDA bf16 logits it runs the library's fused kernel with a freshly drawn Exp(1) tensor."""
    if logits.is_cuda and logits.dtype == torch.bfloat16 and logits.dim() == 2:
        q = torch.empty_like(logits).exponential_(1)
        out = torch.empty(logits.shape[0], dtype=torch.int32, device=logits.device)
        lg = logits.contiguous()
        _native.check(_native.lib().csm_k_sample_topk(
            lg.data_ptr(), q.data_ptr(), lg.shape[0], lg.shape[1], float(temperature), int(topk), out.data_ptr(),
            torch.cuda.current_stream(logits.device).cuda_stream))
        return out.unsqueeze(-1)
    raise RuntimeError("sesameai(B200).sample_topk needs 2-D CUDA bf16 logits (no CPU path)")


class Model(
    nn.Module,
    PyTorchModelHubMixin,
    repo_url="https://github.com/SesameAILabs/csm",
    pipeline_tag="text-to-speech",
    license="apache-2.0",
):
    def __init__(self, config: ModelArgs):
        super().__init__()
        if isinstance(config, dict):
            config = ModelArgs(**config)
        self.config = config
        self.backbone = FLAVORS[config.backbone_flavor]()
        self.decoder = FLAVORS[config.decoder_flavor]()
        d_bb, d_dec = self.backbone.spec.embed_dim, self.decoder.spec.embed_dim
        self.text_embeddings = nn.Embedding(config.text_vocab_size, d_bb)
        self.audio_embeddings = nn.Embedding(config.audio_vocab_size * config.audio_num_codebooks, d_bb)
        self.projection = _Weight(d_dec, d_bb)
        self.codebook0_head = _Weight(config.audio_vocab_size, d_bb)
        self.audio_head = nn.Parameter(torch.empty(config.audio_num_codebooks - 1, d_dec, config.audio_vocab_size))
        self._ctx: Optional[int] = None
        self._keep: Dict[str, object] = {}
        # In-kernel sampling noise is a counter RNG keyed by (seed, frame counter, codebook, stream, index).
        # ``seed`` is re-drawn from torch's default generator by setup_caches() and by every reset_caches()
        # (= per utterance), so torch.manual_seed() governs the audio as it does for the reference (which
        # draws its Exp(1) noise from torch's generator) and unseeded processes / replicas differ.
        # Assign ``model.seed`` (after reset_caches) for a fixed stream in tests.
        self._frame_counter = 0
        self.seed = 0

    # -- life cycle ------------------------------------------------------------------------------
    def _release(self) -> None:
        if getattr(self, "_ctx", None):
            _native.lib().csm_destroy(self._ctx)
        self._ctx = None
        self._keep = {}

    def __del__(self):  # pragma: no cover
        try:
            self._release()
        except Exception:
            pass

    def clone_for_context(self) -> "Model":
        """A second decode context over the SAME parameter tensors: the clone gets its own KV caches, workspace
        and CUDA graphs from its own ``setup_caches`` (``sesameai.serving.LaneGroups`` runs several of them on
        separate CUDA streams).  Nothing is copied except the module's bookkeeping dicts."""
        import copy

        m = copy.copy(self)
        m._parameters = dict(self._parameters)
        m._buffers = dict(self._buffers)
        m._modules = dict(self._modules)
        m._ctx = None
        m._keep = {}
        return m

    def setup_caches(self, max_batch_size: int) -> None:
        """Reference ``Model.setup_caches`` (``models.py:120-130``): allocates the KV caches (inside
        one torch-owned workspace), packs the weights for the kernels and enables generation."""
        p = next(self.parameters())
        if p.device.type != "cuda" or p.dtype != torch.bfloat16:
            raise RuntimeError(
                "sesameai(B200): setup_caches needs the model on a CUDA device in bfloat16 "
                "(model.to(device='cuda', dtype=torch.bfloat16)); there is no CPU path")
        L = _native.lib()
        self._release()
        dev = p.device
        cfg = self._native_config()
        need = L.csm_workspace_bytes(ctypes.byref(cfg), int(max_batch_size))
        if need == 0:
            raise ValueError("sesameai(B200): unsupported model configuration")
        with torch.cuda.device(dev):
            ws = torch.empty(need + 256, dtype=torch.uint8, device=dev)
            off = (-ws.data_ptr()) % 256
            rope_bb = self.backbone.rope_table(self.backbone.max_seq_len).to(dev).contiguous()
            rope_dec = self.decoder.rope_table(self.decoder.max_seq_len).to(dev).contiguous()
            w, keep = self._native_weights(rope_bb, rope_dec)
            ctx = ctypes.c_void_p()
            _native.check(L.csm_create(ctypes.byref(cfg), ctypes.byref(w), int(max_batch_size), ws.data_ptr() + off,
                                       need, torch.cuda.current_stream(dev).cuda_stream, ctypes.byref(ctx)))
        self._ctx = ctx.value
        self._reseed()
        self._keep = {"ws": ws, "rope_bb": rope_bb, "rope_dec": rope_dec, "w": w, "arrays": keep, "max_batch": max_batch_size}
        tri = lambda n: torch.tril(torch.ones(n, n, dtype=torch.bool, device=dev))  # noqa: E731
        self.register_buffer("backbone_causal_mask", tri(self.backbone.max_seq_len))
        self.register_buffer("decoder_causal_mask", tri(self.config.audio_num_codebooks))
        self.backbone.__dict__["_owner_ref"] = weakref.ref(self)
        self.decoder.__dict__["_owner_ref"] = weakref.ref(self)

    def _reseed(self) -> None:
        self.seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        self._frame_counter = 0

    def reset_caches(self) -> None:
        if self._ctx is None:
            raise RuntimeError("Key value caches are not setup. Call ``setup_caches()`` first.")
        _native.check(_native.lib().csm_reset_caches(self._ctx))
        self._reseed()

    def reset_lane(self, lane: int) -> None:
        """Rewind one cache lane (a finished stream leaves; the lane is free for the next request)."""
        if self._ctx is None:
            raise RuntimeError("Key value caches are not setup. Call ``setup_caches()`` first.")
        _native.check(_native.lib().csm_lane_reset(self._ctx, int(lane)))

    def lane_len(self, lane: int) -> int:
        return int(_native.lib().csm_lane_len(self._ctx, int(lane)))

    def check_device_error(self) -> None:
        """Raise for an error that an earlier (stream-ordered, asynchronous) call hit on the device.  Costs a
        host memory read, no CUDA call; meaningful once the stream has been synchronised (the frame loop's
        EOS test does that).  The context stays usable: ``reset_caches()`` and retry."""
        if self._ctx is None:
            return
        code = _native.lib().csm_check_error(self._ctx, 1)
        if code == 0:
            return
        if code == 0x801:
            raise IndexError("sesameai(B200): token id out of range of its embedding table")
        if code == 0x802:
            raise IndexError("sesameai(B200): teacher-forced token id out of range")
        if code == 0x803:
            raise ValueError("sesameai(B200): input_pos must continue the cache position and stay below max_seq_len")
        raise RuntimeError(f"sesameai(B200): decode kernel gave up waiting (code {code:#x}); reset_caches() and retry")

    # -- the hot path ----------------------------------------------------------------------------
    def generate_frame(self, tokens: torch.Tensor, tokens_mask: torch.Tensor, input_pos: torch.Tensor,
                       temperature: float, topk: int, *, noise: Optional[torch.Tensor] = None,
                       forced: Optional[torch.Tensor] = None, logits_out: Optional[torch.Tensor] = None,
                       sampled_out: Optional[torch.Tensor] = None, no_graph: bool = False,
                       path: int = 0, prefill: int = 0, lanes=None) -> torch.Tensor:
        """(B, S, 33) tokens/mask + (B, S) positions -> (B, 32) int32 codes, like the reference
        (``models.py:132-184``).  Keyword extras are for parity tests: shared Exp(1) ``noise``
        [32, B, V] bf16, teacher-``forced`` tokens [B, 32] int32, raw ``logits_out`` [32, B, V].
        ``lanes`` (continuous batching, ``sesameai.serving``): the KV-cache lane of every batch row, distinct
        ints below ``max_batch_size``; each lane keeps its own length, so streams join, advance in any subset and
        leave (``reset_lane``) independently.  Default: row b on lane b, the reference's lock-step batch."""
        assert self._ctx is not None, "backbone caches are not enabled"
        self.check_device_error()  # of earlier calls (free: a host memory read)
        dev = tokens.device
        if dev.type != "cuda":
            raise RuntimeError("sesameai(B200): tokens must live on the model's CUDA device")
        B, S, ncol = tokens.shape
        C = self.config.audio_num_codebooks
        if ncol != C + 1:
            raise ValueError(f"tokens must have {C + 1} columns")
        tok = tokens.to(torch.int64).contiguous()
        msk = tokens_mask.to(torch.bool).contiguous()
        pos = input_pos.to(torch.int64).contiguous()
        out = torch.empty(B, C, dtype=torch.int32, device=dev)
        opts = _native.FrameOpts()
        keep = []
        if noise is not None:
            nz = noise.to(device=dev, dtype=torch.bfloat16).contiguous()
            assert nz.shape == (C, B, self.config.audio_vocab_size)
            keep.append(nz)
            opts.noise = nz.data_ptr()
        else:
            opts.seed = int(self.seed) & (2 ** 64 - 1)
            opts.offset = self._frame_counter
        if forced is not None:
            fz = forced.to(device=dev, dtype=torch.int32).contiguous()
            keep.append(fz)
            opts.forced = fz.data_ptr()
        if logits_out is not None:
            assert logits_out.is_contiguous() and logits_out.dtype == torch.bfloat16
            opts.logits_out = logits_out.data_ptr()
        if sampled_out is not None:
            assert sampled_out.is_contiguous() and sampled_out.dtype == torch.int32
            opts.sampled_out = sampled_out.data_ptr()
        opts.path = _native.PATH_DIRECT if no_graph else int(path)
        opts.prefill = int(prefill)
        if lanes is not None:
            lane_list = [int(v) for v in lanes]
            if len(lane_list) != B:
                raise ValueError("lanes must name one cache lane per batch row")
            lane_arr = (ctypes.c_int32 * B)(*lane_list)
            opts.lanes = lane_arr  # host array, read before the call returns
        self._frame_counter += 1
        with torch.cuda.device(dev):
            rc = _native.lib().csm_generate_frame(
                self._ctx, tok.data_ptr(), msk.data_ptr(), pos.data_ptr(), B, S, float(temperature), int(topk),
                ctypes.byref(opts), out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        if rc == _native.CSM_ERR_OVERFLOW:
            raise AssertionError(_native.lib().csm_last_error().decode())
        if rc == _native.CSM_ERR_STATE:
            raise ValueError(_native.lib().csm_last_error().decode())
        _native.check(rc)
        for t in (tok, msk, pos, *keep):  # keep inputs alive until the stream has consumed them
            t.record_stream(torch.cuda.current_stream(dev))
        return out

    # -- plumbing ----------------------------------------------------------------------------------
    def _native_config(self) -> _native.Config:
        cfg = _native.Config()
        for dst, st in ((cfg.backbone, self.backbone.spec), (cfg.decoder, self.decoder.spec)):
            dst.layers, dst.dim, dst.heads, dst.kv_heads, dst.ff = (
                st.num_layers, st.embed_dim, st.num_heads, st.num_kv_heads, st.intermediate_dim)
        cfg.text_vocab = self.config.text_vocab_size
        cfg.audio_vocab = self.config.audio_vocab_size
        cfg.codebooks = self.config.audio_num_codebooks
        cfg.max_seq_len = self.backbone.max_seq_len
        cfg.norm_eps = self.backbone.spec.norm_eps
        return cfg

    def _native_weights(self, rope_bb: torch.Tensor, rope_dec: torch.Tensor):
        def ptr(t: torch.Tensor) -> int:
            if not t.is_contiguous():
                raise RuntimeError("sesameai(B200): parameters must be contiguous")
            return t.data_ptr()

        def layer_array(stack: TransformerStack):
            arr = (_native.LayerWeights * len(stack.layers))()
            for i, blk in enumerate(stack.layers):
                a = arr[i]
                a.q_proj, a.k_proj = ptr(blk.attn.q_proj.weight), ptr(blk.attn.k_proj.weight)
                a.v_proj, a.output_proj = ptr(blk.attn.v_proj.weight), ptr(blk.attn.output_proj.weight)
                a.w1, a.w2, a.w3 = ptr(blk.mlp.w1.weight), ptr(blk.mlp.w2.weight), ptr(blk.mlp.w3.weight)
                a.sa_norm, a.mlp_norm = ptr(blk.sa_norm.scale), ptr(blk.mlp_norm.scale)
            return arr

        w = _native.Weights()
        w.text_embeddings = ptr(self.text_embeddings.weight)
        w.audio_embeddings = ptr(self.audio_embeddings.weight)
        w.projection = ptr(self.projection.weight)
        w.codebook0_head = ptr(self.codebook0_head.weight)
        w.audio_head = ptr(self.audio_head)
        w.backbone_norm, w.decoder_norm = ptr(self.backbone.norm.scale), ptr(self.decoder.norm.scale)
        w.backbone_rope, w.decoder_rope = rope_bb.data_ptr(), rope_dec.data_ptr()
        w.backbone_rope_len, w.decoder_rope_len = rope_bb.shape[0], rope_dec.shape[0]
        bb, dec = layer_array(self.backbone), layer_array(self.decoder)
        w.backbone_layers, w.decoder_layers = bb, dec
        return w, (bb, dec)
