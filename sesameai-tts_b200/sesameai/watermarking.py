"""Watermarking stays the reference path (BASELINE.json north_star: "outside the timed region").

This module only keeps ``tts_service.py``'s imports working -- ``CSM_1B_GH_WATERMARK``,
``load_watermarker``, ``watermark``, ``verify`` (reference ``sesameai/watermarking.py:9,20-59``) --
by delegating to the third-party ``silentcipher`` package, imported lazily so the frame-generation
hot path has no dependency on it.  Only the two sinc resamplings around the watermarker (24 kHz <-> 44.1 kHz)
run on the B200 (``sesameai.postprocess.resample``) when the audio is a CUDA tensor.
"""
from __future__ import annotations

from typing import List, Tuple

import torch

CSM_1B_GH_WATERMARK = [212, 211, 146, 56, 201]  # public demo key of the upstream project
_WM_RATE = 44_100


def _silentcipher():
    try:
        import silentcipher
    except ImportError as e:  # pragma: no cover
        raise RuntimeError("watermarking needs the `silentcipher` package (reference requirements.txt:9)") from e
    return silentcipher


def _resample(x: torch.Tensor, src: int, dst: int) -> torch.Tensor:
    if src == dst:
        return x
    if x.is_cuda and x.dtype == torch.float32:  # same filter as torchaudio's default resample, on the B200
        from .postprocess import resample

        return resample(x, src, dst)
    import torchaudio

    return torchaudio.functional.resample(x, orig_freq=src, new_freq=dst)


def load_watermarker(device: str = "cuda"):
    return _silentcipher().get_model(model_type="44.1k", device=device)


@torch.inference_mode()
def watermark(watermarker, audio_array: torch.Tensor, sample_rate: int, watermark_key: List[int]) -> Tuple[torch.Tensor, int]:
    marked, _ = watermarker.encode_wav(_resample(audio_array, sample_rate, _WM_RATE), _WM_RATE, watermark_key,
                                       calc_sdr=False, message_sdr=36)
    out_rate = min(_WM_RATE, sample_rate)
    return _resample(marked, _WM_RATE, out_rate), out_rate


@torch.inference_mode()
def verify(watermarker, watermarked_audio: torch.Tensor, sample_rate: int, watermark_key: List[int]) -> bool:
    res = watermarker.decode_wav(_resample(watermarked_audio, sample_rate, _WM_RATE), _WM_RATE, phase_shift_decoding=True)
    return bool(res["status"]) and res["messages"][0] == watermark_key
