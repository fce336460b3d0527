"""ORACLE / TEST INFRASTRUCTURE ONLY -- the product path never imports this file.

CPU/PyTorch restatement of the Mimi codec **decode** path that the reference reaches through
``moshi==0.2.2`` (``requirements.txt:6``; call sites ``sesameai/generator.py:52-57,116,299`` and
``tts_service.py:245``): ``loaders.get_mimi(...)`` -> ``MimiModel.decode(codes[B,32,T])`` ->
``[B,1,1920*T]`` fp32.  moshi is not vendored by the reference and not installed here, so its
published algorithm is restated: split-RVQ dequantisation (1 semantic +
31 acoustic codebooks, 2048 x 256, 1x1 output projections 256->512), depthwise ConvTranspose1d
x2 upsample, 8-layer causal transformer (LayerNorm, packed in_proj, interleaved RoPE
max_period 10000, context 250, LayerScale, GELU MLP), SEANet decoder (ratios 8,6,5,4, ELU,
causal convs, one residual block per stage, true skip).

Pinning: ``tests/test_mimi_oracle_pin_hf.py`` maps these weights onto the independent
``transformers`` ``MimiModel`` port (rotate-half RoPE -> permuted q/k rows) and requires fp32
agreement of the waveform; parameter names follow moshi's state dict so the real
``kyutai/moshiko`` tokenizer checkpoint's decode-side keys would load.
"""
from __future__ import annotations

import math
from typing import List

import torch
import torch.nn.functional as F
from torch import nn

SAMPLE_RATE = 24_000
FRAME_RATE = 12.5
RATIOS = (8, 6, 5, 4)
DIM = 512
N_FILTERS = 64
N_Q = 32
BINS = 2048
Q_DIM = 256
TR_LAYERS, TR_HEADS, TR_FF, TR_CONTEXT, TR_MAX_PERIOD = 8, 8, 2048, 250, 10_000.0


def causal_conv1d(x: torch.Tensor, weight: torch.Tensor, bias, dilation: int = 1) -> torch.Tensor:
    """moshi StreamingConv1d (causal, stride 1, pad_mode constant): left-pad (k-1)*dilation zeros."""
    k = weight.shape[-1]
    return F.conv1d(F.pad(x, ((k - 1) * dilation, 0)), weight, bias, dilation=dilation)


def causal_convtr1d(x: torch.Tensor, weight: torch.Tensor, bias, stride: int, groups: int = 1) -> torch.Tensor:
    """moshi StreamingConvTranspose1d (causal, trim_right_ratio 1): trim kernel-stride samples on the right."""
    y = F.conv_transpose1d(x, weight, bias, stride=stride, groups=groups)
    trim = weight.shape[-1] - stride
    return y[..., : y.shape[-1] - trim]


class _Codebook(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer("_initialized", torch.tensor([1.0]))
        self.register_buffer("cluster_usage", torch.ones(BINS))
        self.register_buffer("embedding_sum", torch.zeros(BINS, Q_DIM))

    @property
    def embedding(self) -> torch.Tensor:
        return self.embedding_sum / self.cluster_usage.clamp(min=1e-5)[:, None]


class _VQ(nn.Module):
    def __init__(self):
        super().__init__()
        self._codebook = _Codebook()


class _RVQ(nn.Module):
    def __init__(self, n_q: int):
        super().__init__()
        self.vq = nn.Module()
        self.vq.layers = nn.ModuleList([_VQ() for _ in range(n_q)])
        self.input_proj = nn.Conv1d(DIM, Q_DIM, 1, bias=False)
        self.output_proj = nn.Conv1d(Q_DIM, DIM, 1, bias=False)

    def decode(self, codes: torch.Tensor) -> torch.Tensor:  # [B, K, T] -> [B, 512, T]
        q = 0.0
        for k in range(codes.shape[1]):
            q = q + F.embedding(codes[:, k], self.vq.layers[k]._codebook.embedding)  # [B, T, 256]
        return self.output_proj(q.transpose(1, 2))


class _TLayer(nn.Module):
    def __init__(self):
        super().__init__()
        self.self_attn = nn.Module()
        self.self_attn.in_proj_weight = nn.Parameter(torch.empty(3 * DIM, DIM))
        self.self_attn.out_proj = nn.Linear(DIM, DIM, bias=False)
        self.norm1 = nn.LayerNorm(DIM, eps=1e-5)
        self.norm2 = nn.LayerNorm(DIM, eps=1e-5)
        self.linear1 = nn.Linear(DIM, TR_FF, bias=False)
        self.linear2 = nn.Linear(TR_FF, DIM, bias=False)
        self.layer_scale_1 = nn.Module()
        self.layer_scale_1.scale = nn.Parameter(torch.full((DIM,), 0.01))
        self.layer_scale_2 = nn.Module()
        self.layer_scale_2.scale = nn.Parameter(torch.full((DIM,), 0.01))


def _rope_interleaved(q: torch.Tensor, k: torch.Tensor):
    """moshi ``apply_rope``: pairs (2i, 2i+1), freq_i = exp(-ln(max_period) * 2i / D), fp32."""
    B, H, T, D = q.shape
    ds = torch.arange(D // 2, dtype=torch.float32)
    freqs = torch.exp(ds * (-math.log(TR_MAX_PERIOD) * 2 / D))
    ts = torch.arange(T, dtype=torch.float32).view(-1, 1)
    rotr, roti = torch.cos(freqs * ts), torch.sin(freqs * ts)

    def rot(x):
        x = x.view(B, H, T, D // 2, 2)
        xr, xi = x[..., 0].float(), x[..., 1].float()
        return torch.stack([xr * rotr - xi * roti, xr * roti + xi * rotr], dim=-1).view(B, H, T, D).to(x.dtype)

    return rot(q), rot(k)


class OracleMimi(nn.Module):
    """Decode half of moshi's ``MimiModel`` (state-dict names follow moshi)."""

    sample_rate = SAMPLE_RATE
    frame_rate = FRAME_RATE

    def __init__(self):
        super().__init__()
        self.quantizer = nn.Module()
        self.quantizer.rvq_first = _RVQ(1)
        self.quantizer.rvq_rest = _RVQ(N_Q - 1)
        self.upsample = nn.Module()
        self.upsample.convtr = nn.Module()
        self.upsample.convtr.convtr = nn.Module()
        self.upsample.convtr.convtr.convtr = nn.ConvTranspose1d(DIM, DIM, 4, stride=2, groups=DIM, bias=False)
        self.decoder_transformer = nn.Module()
        self.decoder_transformer.transformer = nn.Module()
        self.decoder_transformer.transformer.layers = nn.ModuleList([_TLayer() for _ in range(TR_LAYERS)])
        # SEANet decoder: model.0 conv, then per ratio [ELU, convtr, resblock], ELU, final conv
        self.decoder = nn.Module()
        model: List[nn.Module] = []

        def conv(cin, cout, k):
            m = nn.Module()
            m.conv = nn.Module()
            m.conv.conv = nn.Conv1d(cin, cout, k)
            return m

        def convtr(cin, cout, k, s):
            m = nn.Module()
            m.convtr = nn.Module()
            m.convtr.convtr = nn.ConvTranspose1d(cin, cout, k, stride=s)
            return m

        ch = N_FILTERS * 2 ** len(RATIOS)
        model.append(conv(DIM, ch, 7))
        for r in RATIOS:
            model.append(nn.ELU())
            model.append(convtr(ch, ch // 2, 2 * r, r))
            res = nn.Module()
            res.block = nn.ModuleList([nn.ELU(), conv(ch // 2, ch // 4, 3), nn.ELU(), conv(ch // 4, ch // 2, 1)])
            model.append(res)
            ch //= 2
        model.append(nn.ELU())
        model.append(conv(N_FILTERS, 1, 3))
        self.decoder.model = nn.ModuleList(model)
        self.num_codebooks = N_Q
        self._build_encoder()

    def set_num_codebooks(self, n: int) -> None:
        self.num_codebooks = n

    # -- stages ----------------------------------------------------------------------------------
    def dequantize(self, codes: torch.Tensor) -> torch.Tensor:
        return self.quantizer.rvq_first.decode(codes[:, :1]) + self.quantizer.rvq_rest.decode(codes[:, 1:])

    def upsample_2x(self, x: torch.Tensor) -> torch.Tensor:
        return causal_convtr1d(x, self.upsample.convtr.convtr.convtr.weight, None, stride=2, groups=DIM)

    def transformer(self, x: torch.Tensor) -> torch.Tensor:  # [B, 512, T'] -> same (conv layout)
        h = x.transpose(1, 2)
        B, T, _ = h.shape
        pos = torch.arange(T)
        delta = pos.view(-1, 1) - pos.view(1, -1)
        allowed = (delta >= 0) & (delta < TR_CONTEXT)
        for layer in self.decoder_transformer.transformer.layers:
            n = layer.norm1(h)
            qkv = F.linear(n, layer.self_attn.in_proj_weight).view(B, T, 3, TR_HEADS, DIM // TR_HEADS)
            q, k, v = (qkv[:, :, i].transpose(1, 2) for i in range(3))
            q, k = _rope_interleaved(q, k)
            a = F.scaled_dot_product_attention(q, k, v, attn_mask=allowed)
            a = layer.self_attn.out_proj(a.transpose(1, 2).reshape(B, T, DIM))
            h = h + layer.layer_scale_1.scale * a
            m = layer.linear2(F.gelu(layer.linear1(layer.norm2(h))))
            h = h + layer.layer_scale_2.scale * m
        return h.transpose(1, 2)

    def seanet(self, x: torch.Tensor) -> torch.Tensor:
        mods = self.decoder.model
        i = 0
        x = causal_conv1d(x, mods[0].conv.conv.weight, mods[0].conv.conv.bias)
        i = 1
        for r in RATIOS:
            x = F.elu(x)
            ct = mods[i + 1].convtr.convtr
            x = causal_convtr1d(x, ct.weight, ct.bias, stride=r)
            blk = mods[i + 2].block
            y = causal_conv1d(F.elu(x), blk[1].conv.conv.weight, blk[1].conv.conv.bias)
            y = causal_conv1d(F.elu(y), blk[3].conv.conv.weight, blk[3].conv.conv.bias)
            x = x + y
            i += 3
        x = F.elu(x)
        return causal_conv1d(x, mods[i + 1].conv.conv.weight, mods[i + 1].conv.conv.bias)

    @torch.no_grad()
    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        """codes [B, K<=32, T] int -> waveform [B, 1, 1920*T] fp32 (moshi ``MimiModel.decode``)."""
        codes = codes.long()
        emb = self.dequantize(codes)
        emb = self.upsample_2x(emb)
        emb = self.transformer(emb)
        return self.seanet(emb)

    # -- encode (reference call site sesameai/generator.py:86) ---------------
    def _build_encoder(self) -> None:
        """SEANet encoder (ratios reversed 4,5,6,8), encoder transformer, stride-2 downsample."""

        def conv(cin, cout, k, stride=1, bias=True):
            m = nn.Module()
            m.conv = nn.Module()
            m.conv.conv = nn.Conv1d(cin, cout, k, stride=stride, bias=bias)
            return m

        model: List[nn.Module] = [conv(1, N_FILTERS, 7)]
        ch = N_FILTERS
        for r in reversed(RATIOS):
            res = nn.Module()
            res.block = nn.ModuleList([nn.ELU(), conv(ch, ch // 2, 3), nn.ELU(), conv(ch // 2, ch, 1)])
            model += [res, nn.ELU(), conv(ch, 2 * ch, 2 * r, stride=r)]
            ch *= 2
        model += [nn.ELU(), conv(ch, DIM, 3)]
        self.encoder = nn.Module()
        self.encoder.model = nn.ModuleList(model)
        self.encoder_transformer = nn.Module()
        self.encoder_transformer.transformer = nn.Module()
        self.encoder_transformer.transformer.layers = nn.ModuleList([_TLayer() for _ in range(TR_LAYERS)])
        self.downsample = nn.Module()
        self.downsample.conv = nn.Module()
        self.downsample.conv.conv = nn.Module()
        self.downsample.conv.conv.conv = nn.Conv1d(DIM, DIM, 4, stride=2, bias=False)

    def seanet_encode(self, x: torch.Tensor) -> torch.Tensor:
        mods = self.encoder.model

        def sconv(x, m, stride=1):  # causal strided conv: left pad k - stride zeros
            c = m.conv.conv
            return F.conv1d(F.pad(x, (c.kernel_size[0] - stride, 0)), c.weight, c.bias, stride=stride)

        x = sconv(x, mods[0])
        i = 1
        for r in reversed(RATIOS):
            blk = mods[i].block
            y = sconv(F.elu(x), blk[1])
            y = sconv(F.elu(y), blk[3])
            x = x + y
            x = sconv(F.elu(x), mods[i + 2], stride=r)
            i += 3
        return sconv(F.elu(x), mods[i + 1])

    def _transformer_layers(self, x: torch.Tensor, layers) -> torch.Tensor:
        saved = self.decoder_transformer.transformer.layers
        self.decoder_transformer.transformer.layers = layers
        try:
            return self.transformer(x)
        finally:
            self.decoder_transformer.transformer.layers = saved

    @staticmethod
    def _rvq_encode(rvq: "_RVQ", x: torch.Tensor, n_q: int) -> torch.Tensor:
        """moshi ResidualVectorQuantization.encode: nearest centroid (cdist + argmin) per layer on the
        running residual of the input projection."""
        res = rvq.input_proj(x).transpose(1, 2)  # [B, T, 256]
        out = []
        for k in range(n_q):
            emb = rvq.vq.layers[k]._codebook.embedding
            idx = torch.cdist(res.reshape(1, -1, Q_DIM), emb[None], p=2)[0].argmin(dim=-1).view(res.shape[:2])
            out.append(idx)
            res = res - F.embedding(idx, emb)
        return torch.stack(out, dim=1)  # [B, n_q, T]

    @torch.no_grad()
    def encode_latent(self, wav: torch.Tensor) -> torch.Tensor:
        """wav [B, 1, L] -> the 12.5 Hz latent [B, 512, T] the split RVQ quantises (everything of ``encode``
        before the search)."""
        if not hasattr(self, "encoder"):
            raise RuntimeError("call _build_encoder() before loading encoder weights")
        L = wav.shape[-1]
        pad = (-L) % 1920
        x = F.pad(wav, (0, pad)) if pad else wav
        emb = self.seanet_encode(x)
        emb = self._transformer_layers(emb, self.encoder_transformer.transformer.layers)
        ds = self.downsample.conv.conv.conv
        return F.conv1d(F.pad(emb, (2, 0), mode="replicate"), ds.weight, None, stride=2)

    @torch.no_grad()
    def quantize_latent(self, emb: torch.Tensor) -> torch.Tensor:
        """latent [B, 512, T] -> codes [B, num_codebooks, T] (moshi SplitResidualVectorQuantizer.encode)."""
        nq = self.num_codebooks
        first = self._rvq_encode(self.quantizer.rvq_first, emb, 1)
        if nq > 1:
            return torch.cat([first, self._rvq_encode(self.quantizer.rvq_rest, emb, nq - 1)], dim=1)
        return first

    @torch.no_grad()
    def latent_for_codes(self, codes: torch.Tensor) -> torch.Tensor:
        """Test helper (not a moshi function): a latent [B, 512, T] whose two input projections land exactly on
        the centroid sums of ``codes`` [B, K, T] -- least-squares solve of [W_first; W_rest] z = [e_0 ; sum_k e_k]
        in fp64 -- so the search's decisions have the margins of the codebooks, not of encoder rounding."""
        B, K, T = codes.shape
        wf = self.quantizer.rvq_first.input_proj.weight[:, :, 0].double()
        wr = self.quantizer.rvq_rest.input_proj.weight[:, :, 0].double()
        tf = self.quantizer.rvq_first.vq.layers[0]._codebook.embedding.double()[codes[:, 0]]  # [B, T, 256]
        tr = torch.zeros_like(tf)
        for k in range(1, K):
            tr = tr + self.quantizer.rvq_rest.vq.layers[k - 1]._codebook.embedding.double()[codes[:, k]]
        A = torch.cat([wf, wr], dim=0)                      # [512, 512]
        rhs = torch.cat([tf, tr], dim=-1).reshape(-1, 512)  # [B*T, 512]
        z = torch.linalg.lstsq(A, rhs.t()).solution.t()     # [B*T, 512]
        return z.reshape(B, T, 512).transpose(1, 2).float().contiguous()

    @torch.no_grad()
    def encode(self, wav: torch.Tensor) -> torch.Tensor:
        """wav [B, 1, L] fp32 at 24 kHz -> codes [B, num_codebooks, ceil(L / 1920)] int64 (moshi
        ``MimiModel.encode``; the waveform is zero-padded on the right to a whole number of frames)."""
        return self.quantize_latent(self.encode_latent(wav))
