"""Mimi codec, decode side, for B200: stands where moshi's ``MimiModel`` stands in the reference
(``loaders.get_mimi`` at reference ``sesameai/generator.py:52-57``; ``.decode`` at ``:116,299`` and
``tts_service.py:245``).  Parameters keep moshi's state-dict names so the decode-side tensors of
the ``kyutai/moshiko`` tokenizer checkpoint load as they are; the arithmetic runs in
libcsm_b200.so (``mimi_decode``).  No PyTorch / CPU execution path exists here.

``encode`` (voice-prompt audio -> codes) runs in the same library
(``mimi_encode``): SEANet encoder, encoder transformer, stride-2 downsample, split-RVQ search.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional

import torch
from torch import nn

from . import _native

SAMPLE_RATE = 24_000
FRAME_RATE = 12.5
DEFAULT_REPO = "kyutai/moshiko-pytorch-bf16"
MIMI_NAME = "tokenizer-e351c8d8-checkpoint125.safetensors"
_RATIOS = (8, 6, 5, 4)


def _conv(cin: int, cout: int, k: int, bias: bool = True) -> nn.Module:
    m = nn.Module()
    m.conv = nn.Module()
    m.conv.conv = nn.Module()
    m.conv.conv.weight = nn.Parameter(torch.empty(cout, cin, k))
    if bias:
        m.conv.conv.bias = nn.Parameter(torch.empty(cout))
    return m


def _transformer_layers() -> nn.ModuleList:
    layers = []
    for _ in range(8):
        l = nn.Module()
        l.self_attn = nn.Module()
        l.self_attn.in_proj_weight = nn.Parameter(torch.empty(1536, 512))
        l.self_attn.out_proj = nn.Module()
        l.self_attn.out_proj.weight = nn.Parameter(torch.empty(512, 512))
        for nm in ("norm1", "norm2"):
            n = nn.Module()
            n.weight = nn.Parameter(torch.ones(512))
            n.bias = nn.Parameter(torch.zeros(512))
            setattr(l, nm, n)
        l.linear1 = nn.Module()
        l.linear1.weight = nn.Parameter(torch.empty(2048, 512))
        l.linear2 = nn.Module()
        l.linear2.weight = nn.Parameter(torch.empty(512, 2048))
        for nm in ("layer_scale_1", "layer_scale_2"):
            sc = nn.Module()
            sc.scale = nn.Parameter(torch.full((512,), 0.01))
            setattr(l, nm, sc)
        layers.append(l)
    return nn.ModuleList(layers)


def _convtr(cin: int, cout: int, k: int) -> nn.Module:
    m = nn.Module()
    m.convtr = nn.Module()
    m.convtr.convtr = nn.Module()
    m.convtr.convtr.weight = nn.Parameter(torch.empty(cin, cout, k))
    m.convtr.convtr.bias = nn.Parameter(torch.empty(cout))
    return m


class _RVQ(nn.Module):
    def __init__(self, n_q: int):
        super().__init__()
        self.vq = nn.Module()
        layers = []
        for _ in range(n_q):
            l = nn.Module()
            l._codebook = nn.Module()
            l._codebook.register_buffer("_initialized", torch.ones(1))
            l._codebook.register_buffer("cluster_usage", torch.ones(2048))
            l._codebook.register_buffer("embedding_sum", torch.zeros(2048, 256))
            layers.append(l)
        self.vq.layers = nn.ModuleList(layers)
        self.input_proj = nn.Module()
        self.input_proj.weight = nn.Parameter(torch.empty(256, 512, 1))
        self.output_proj = nn.Module()
        self.output_proj.weight = nn.Parameter(torch.empty(512, 256, 1))


class MimiCodec(nn.Module):
    """Duck-types the part of moshi's ``MimiModel`` the reference uses: ``decode``, ``sample_rate``,
    ``frame_rate``, ``set_num_codebooks``."""

    sample_rate = SAMPLE_RATE
    frame_rate = FRAME_RATE

    def __init__(self, max_frames: int = 1200):
        super().__init__()
        self.max_frames = max_frames
        self.num_codebooks = 32
        self.quantizer = nn.Module()
        self.quantizer.rvq_first = _RVQ(1)
        self.quantizer.rvq_rest = _RVQ(31)
        self.upsample = nn.Module()
        self.upsample.convtr = nn.Module()
        self.upsample.convtr.convtr = nn.Module()
        self.upsample.convtr.convtr.convtr = nn.Module()
        self.upsample.convtr.convtr.convtr.weight = nn.Parameter(torch.empty(512, 1, 4))
        self.decoder_transformer = nn.Module()
        self.decoder_transformer.transformer = nn.Module()
        self.decoder_transformer.transformer.layers = _transformer_layers()
        self.decoder = nn.Module()
        model: List[nn.Module] = [_conv(512, 1024, 7)]
        ch = 1024
        for r in _RATIOS:
            model.append(nn.Identity())  # ELU slot
            model.append(_convtr(ch, ch // 2, 2 * r))
            res = nn.Module()
            res.block = nn.ModuleList([nn.Identity(), _conv(ch // 2, ch // 4, 3), nn.Identity(), _conv(ch // 4, ch // 2, 1)])
            model.append(res)
            ch //= 2
        model.append(nn.Identity())
        model.append(_conv(64, 1, 3))
        self.decoder.model = nn.ModuleList(model)
        # encode side: SEANet encoder (ratios reversed), encoder transformer, stride-2 downsample
        enc: List[nn.Module] = [_conv(1, 64, 7)]
        ch = 64
        for r in reversed(_RATIOS):
            res = nn.Module()
            res.block = nn.ModuleList([nn.Identity(), _conv(ch, ch // 2, 3), nn.Identity(), _conv(ch // 2, ch, 1)])
            enc += [res, nn.Identity(), _conv(ch, 2 * ch, 2 * r)]
            ch *= 2
        enc += [nn.Identity(), _conv(ch, 512, 3)]
        self.encoder = nn.Module()
        self.encoder.model = nn.ModuleList(enc)
        self.encoder_transformer = nn.Module()
        self.encoder_transformer.transformer = nn.Module()
        self.encoder_transformer.transformer.layers = _transformer_layers()
        self.downsample = nn.Module()
        self.downsample.conv = _conv(512, 512, 4, bias=False)
        self._ctx: Optional[int] = None
        self._keep: Dict[str, object] = {}

    def set_num_codebooks(self, n: int) -> None:
        if not 1 <= n <= 32:
            raise ValueError("num_codebooks must be in [1, 32]")
        self.num_codebooks = n

    # ------------------------------------------------------------------------------------------
    def _weight_list(self) -> List[torch.Tensor]:
        sd = self.state_dict()
        w: List[torch.Tensor] = []
        for k in range(32):
            pre = "quantizer.rvq_first.vq.layers.0." if k == 0 else f"quantizer.rvq_rest.vq.layers.{k - 1}."
            w += [sd[pre + "_codebook.embedding_sum"], sd[pre + "_codebook.cluster_usage"]]
        w += [sd["quantizer.rvq_first.output_proj.weight"], sd["quantizer.rvq_rest.output_proj.weight"],
              sd["upsample.convtr.convtr.convtr.weight"]]
        for l in range(8):
            pre = f"decoder_transformer.transformer.layers.{l}."
            w += [sd[pre + n] for n in ("self_attn.in_proj_weight", "self_attn.out_proj.weight", "norm1.weight", "norm1.bias",
                                        "norm2.weight", "norm2.bias", "linear1.weight", "linear2.weight",
                                        "layer_scale_1.scale", "layer_scale_2.scale")]
        w += [sd["decoder.model.0.conv.conv.weight"], sd["decoder.model.0.conv.conv.bias"]]
        for s in range(4):
            i = 2 + 3 * s
            w += [sd[f"decoder.model.{i}.convtr.convtr.weight"], sd[f"decoder.model.{i}.convtr.convtr.bias"],
                  sd[f"decoder.model.{i + 1}.block.1.conv.conv.weight"], sd[f"decoder.model.{i + 1}.block.1.conv.conv.bias"],
                  sd[f"decoder.model.{i + 1}.block.3.conv.conv.weight"], sd[f"decoder.model.{i + 1}.block.3.conv.conv.bias"]]
        w += [sd["decoder.model.14.conv.conv.weight"], sd["decoder.model.14.conv.conv.bias"]]
        # encode side, in MIMI_W_ENC_* order
        w += [sd["encoder.model.0.conv.conv.weight"], sd["encoder.model.0.conv.conv.bias"]]
        for s in range(4):
            i = 1 + 3 * s
            w += [sd[f"encoder.model.{i}.block.1.conv.conv.weight"], sd[f"encoder.model.{i}.block.1.conv.conv.bias"],
                  sd[f"encoder.model.{i}.block.3.conv.conv.weight"], sd[f"encoder.model.{i}.block.3.conv.conv.bias"],
                  sd[f"encoder.model.{i + 2}.conv.conv.weight"], sd[f"encoder.model.{i + 2}.conv.conv.bias"]]
        w += [sd["encoder.model.14.conv.conv.weight"], sd["encoder.model.14.conv.conv.bias"]]
        for l in range(8):
            pre = f"encoder_transformer.transformer.layers.{l}."
            w += [sd[pre + n] for n in ("self_attn.in_proj_weight", "self_attn.out_proj.weight", "norm1.weight", "norm1.bias",
                                        "norm2.weight", "norm2.bias", "linear1.weight", "linear2.weight",
                                        "layer_scale_1.scale", "layer_scale_2.scale")]
        w += [sd["downsample.conv.conv.conv.weight"], sd["quantizer.rvq_first.input_proj.weight"],
              sd["quantizer.rvq_rest.input_proj.weight"]]
        assert len(w) == _native.MIMI_W_COUNT
        return w

    def _release(self) -> None:
        if getattr(self, "_ctx", None):
            _native.lib().mimi_destroy(self._ctx)
        self._ctx = None
        self._keep = {}

    def __del__(self):  # pragma: no cover
        try:
            self._release()
        except Exception:
            pass

    def prepare(self) -> None:
        """Pack the weights for the kernels (call again after loading new weights)."""
        p = next(self.parameters())
        if p.device.type != "cuda" or p.dtype != torch.float32:
            raise RuntimeError("sesameai(B200): the Mimi codec must be on a CUDA device in float32 (no CPU path)")
        L = _native.lib()
        self._release()
        ws_list = [t.detach().contiguous() for t in self._weight_list()]
        arr = (ctypes.c_void_p * len(ws_list))(*[t.data_ptr() for t in ws_list])
        need = L.mimi_workspace_bytes(self.max_frames)
        with torch.cuda.device(p.device):
            ws = torch.empty(need + 256, dtype=torch.uint8, device=p.device)
            off = (-ws.data_ptr()) % 256
            ctx = ctypes.c_void_p()
            _native.check(L.mimi_create(arr, len(ws_list), self.max_frames, ws.data_ptr() + off, need,
                                        torch.cuda.current_stream(p.device).cuda_stream, ctypes.byref(ctx)))
        self._ctx = ctx.value
        self._keep = {"ws": ws, "weights": ws_list}

    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        """codes [B, K, T] (any integer dtype) -> [B, 1, 1920*T] fp32, as moshi's ``MimiModel.decode``."""
        if codes.dim() != 3:
            raise ValueError("codes must be [B, K, T]")
        if self._ctx is None:
            self.prepare()
        dev = next(self.parameters()).device
        c = codes.to(device=dev, dtype=torch.int64).contiguous()
        B, K, T = c.shape
        out = torch.empty(B, 1, 1920 * T, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _native.check(_native.lib().mimi_decode(self._ctx, c.data_ptr(), B, K, T, out.data_ptr(),
                                                    torch.cuda.current_stream(dev).cuda_stream))
        c.record_stream(torch.cuda.current_stream(dev))
        return out

    def quantize_latent(self, latent: torch.Tensor) -> torch.Tensor:
        """latent [1, 512, T] fp32 -> codes [1, num_codebooks, T]: the split-RVQ search of ``encode`` alone
        (moshi ``SplitResidualVectorQuantizer.encode``); used by the parity tests."""
        if latent.dim() != 3 or latent.shape[0] != 1 or latent.shape[1] != 512:
            raise ValueError("latent must be [1, 512, T]")
        if self._ctx is None:
            self.prepare()
        dev = next(self.parameters()).device
        x = latent[0].to(device=dev, dtype=torch.float32).t().contiguous()  # time-major [T, 512]
        T = x.shape[0]
        codes = torch.empty(1, self.num_codebooks, T, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _native.check(_native.lib().mimi_k_rvq_encode(self._ctx, x.data_ptr(), T, self.num_codebooks, codes.data_ptr(),
                                                         torch.cuda.current_stream(dev).cuda_stream))
        x.record_stream(torch.cuda.current_stream(dev))
        return codes

    def streaming(self) -> "MimiStream":
        """A stateful decoder for ONE utterance delivered in chunks (``generate_stream``): concatenated chunk
        outputs equal ``decode`` of the whole utterance."""
        if self._ctx is None:
            self.prepare()
        return MimiStream(self)

    def encode(self, wav: torch.Tensor) -> torch.Tensor:
        """wav [B, 1, L] fp32 at 24 kHz -> codes [B, num_codebooks, ceil(L/1920)] int64, as moshi's
        ``MimiModel.encode`` (reference ``generator.py:86``)."""
        if wav.dim() != 3 or wav.shape[1] != 1:
            raise ValueError("wav must be [B, 1, L]")
        dev = next(self.parameters()).device
        w = wav.to(device=dev, dtype=torch.float32).contiguous()
        B, _, L = w.shape
        T = (L + 1919) // 1920
        if T > self.max_frames:  # moshi has no length limit: grow the workspace (decode windows instead)
            self.max_frames = 1 << (T - 1).bit_length()
            self._release()
        if self._ctx is None:
            self.prepare()
        codes = torch.empty(B, self.num_codebooks, T, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _native.check(_native.lib().mimi_encode(self._ctx, w.data_ptr(), B, L, self.num_codebooks, codes.data_ptr(),
                                                    torch.cuda.current_stream(dev).cuda_stream))
        w.record_stream(torch.cuda.current_stream(dev))
        return codes


class MimiStream:
    """Streaming decode state of one utterance (``mimi_stream`` in include/csm_b200.h)."""

    def __init__(self, codec: MimiCodec):
        L = _native.lib()
        self.codec = codec
        dev = next(codec.parameters()).device
        self.device = dev
        need = L.mimi_stream_state_bytes()
        with torch.cuda.device(dev):
            self._buf = torch.empty(need + 256, dtype=torch.uint8, device=dev)
            off = (-self._buf.data_ptr()) % 256
            h = ctypes.c_void_p()
            _native.check(L.mimi_stream_create(codec._ctx, self._buf.data_ptr() + off, need,
                                               torch.cuda.current_stream(dev).cuda_stream, ctypes.byref(h)))
        self._h = h.value

    def __del__(self):  # pragma: no cover
        try:
            if getattr(self, "_h", None):
                _native.lib().mimi_stream_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def reset(self) -> None:
        with torch.cuda.device(self.device):
            _native.check(_native.lib().mimi_stream_reset(self._h, torch.cuda.current_stream(self.device).cuda_stream))

    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        """The next frames of the utterance: codes [1, K, T] (or [K, T]) -> [1, 1, 1920*T] fp32."""
        if codes.dim() == 3:
            if codes.shape[0] != 1:
                raise ValueError("a MimiStream decodes one utterance")
            codes = codes[0]
        c = codes.to(device=self.device, dtype=torch.int64).contiguous()
        K, T = c.shape
        out = torch.empty(1, 1, 1920 * T, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _native.check(_native.lib().mimi_decode_stream(self._h, c.data_ptr(), K, T, out.data_ptr(),
                                                          torch.cuda.current_stream(self.device).cuda_stream))
        c.record_stream(torch.cuda.current_stream(self.device))
        return out


def get_mimi(filename: Optional[str], device="cuda", max_frames: int = 1200) -> MimiCodec:
    """Counterpart of moshi ``loaders.get_mimi``: build the codec and load a safetensors checkpoint."""
    codec = MimiCodec(max_frames=max_frames)
    if filename is not None:
        from safetensors.torch import load_file

        sd = load_file(filename)
        own = codec.state_dict()
        missing = [k for k in own if k not in sd]
        if missing:
            raise RuntimeError(f"Mimi checkpoint lacks tensors: {missing[:4]} ...")
        codec.load_state_dict({k: sd[k].float() for k in own})
    codec.to(device=device, dtype=torch.float32)
    return codec
