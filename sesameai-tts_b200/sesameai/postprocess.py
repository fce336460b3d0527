"""Waveform post-processing on the B200: the steps the reference runs AFTER Mimi decode
with torchaudio / numpy / pydub on the host -- sinc resampling around the watermarker
(``sesameai/watermarking.py:35-39``, ``tts_service.py:254-256``) and ``generate_audio_segment``'s peak
normalisation, 16-bit conversion, silence padding and fades (``tts_service.py:287-306``) -- as calls into
libcsm_b200.so.  CUDA tensors in, CUDA tensors out; no CPU path."""
from __future__ import annotations

import torch

from . import _native


def resample(waveform: torch.Tensor, orig_freq: int, new_freq: int) -> torch.Tensor:
    """``torchaudio.functional.resample(waveform, orig_freq, new_freq)`` (default arguments) for a 1-D or
    [..., time] CUDA fp32 tensor."""
    if not waveform.is_cuda:
        raise RuntimeError("sesameai(B200).postprocess.resample needs a CUDA tensor (no CPU path)")
    if orig_freq == new_freq:
        return waveform
    L = _native.lib()
    x = waveform.to(torch.float32).contiguous()
    lead, n = x.shape[:-1], x.shape[-1]
    rows = x.reshape(-1, n)
    n_out = L.csm_post_resample_len(n, int(orig_freq), int(new_freq))
    if n_out < 0:
        raise ValueError("unsupported sample rates")
    dev = x.device
    out = torch.empty(rows.shape[0], n_out, dtype=torch.float32, device=dev)
    wsb = L.csm_post_resample_workspace_bytes(int(orig_freq), int(new_freq))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        for r in range(rows.shape[0]):
            _native.check(L.csm_post_resample(rows[r].data_ptr(), n, int(orig_freq), int(new_freq), out[r].data_ptr(),
                                              ws.data_ptr(), wsb, st))
    return out.reshape(*lead, n_out)


def pcm16_segment(audio: torch.Tensor, sample_rate: int, fade_duration: int = 50, start_silence_duration: int = 500,
                  end_silence_duration: int = 100) -> torch.Tensor:
    """``TTS.generate_audio_segment`` after generation (``tts_service.py:287-306``): normalise to the peak, convert to
    int16, add silence (ms) in front / behind, fade in and out (ms; pydub's precise per-sample fades, <= 100 ms).
    Returns the int16 samples of the finished segment (what ``AudioSegment.raw_data`` holds) on the device."""
    if not audio.is_cuda:
        raise RuntimeError("sesameai(B200).postprocess.pcm16_segment needs a CUDA tensor (no CPU path)")
    if fade_duration > 100:
        raise ValueError("fades longer than 100 ms use pydub's coarse per-millisecond form, which is not implemented")
    x = audio.to(torch.float32).reshape(-1).contiguous()
    n = x.numel()
    ms = lambda d: int(d * (sample_rate / 1000.0))  # noqa: E731  (pydub frame_count(ms=...))
    s0, s1, f = ms(start_silence_duration), ms(end_silence_duration), ms(fade_duration)
    dev = x.device
    out = torch.empty(s0 + n + s1, dtype=torch.int16, device=dev)
    scratch = torch.empty(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _native.check(_native.lib().csm_post_pcm16_segment(x.data_ptr(), n, s0, s1, f, f, out.data_ptr(), scratch.data_ptr(),
                                                           torch.cuda.current_stream(dev).cuda_stream))
    return out
