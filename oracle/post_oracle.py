"""CPU restatement of the post-processing steps (TEST INFRASTRUCTURE -- only tests/ may import this).

* ``resample``: torchaudio 2.x ``functional.resample`` with default arguments (``sinc_interp_hann``,
  ``lowpass_filter_width=6``, ``rolloff=0.99``), restated from its published algorithm
  (``_get_sinc_resample_kernel`` / ``_apply_sinc_resample_kernel``).  torchaudio is NOT installed in this image:
  this part is **parity unpinned** (checked only through properties: identity at equal rates, DC gain, tone
  frequency preservation, round trip).  Reference call sites: sesameai/watermarking.py:35-39, tts_service.py:254-256.
* ``pcm16_segment``: tts_service.py:287-306 (numpy normalise / astype int16) + pydub ``AudioSegment.silent``,
  ``+``, ``fade_in`` / ``fade_out`` for fades <= 100 ms, whose arithmetic is CPython's ``audioop.mul``
  (pydub/audio_segment.py ``fade``) -- pinned against the real ``audioop`` of this interpreter in the tests.
"""
import math

import numpy as np
import torch


def resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    g = math.gcd(int(orig_freq), int(new_freq))
    of, nf = int(orig_freq) // g, int(new_freq) // g
    base = min(of, nf) * rolloff
    width = math.ceil(lowpass_filter_width * of / base)
    idx = torch.arange(-width, width + of, dtype=torch.float64)[None, None] / of
    t = torch.arange(0, -nf, -1, dtype=torch.float64)[:, None, None] / nf + idx
    t = t * base
    t = t.clamp(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    scale = base / of
    kernels = torch.where(t == 0, torch.tensor(1.0, dtype=torch.float64), t.sin() / t)
    kernels = kernels * window * scale
    return kernels.to(torch.float32), width, of, nf


def resample(waveform: torch.Tensor, orig_freq: int, new_freq: int) -> torch.Tensor:
    if orig_freq == new_freq:
        return waveform
    kernel, width, of, nf = resample_kernel(orig_freq, new_freq)
    shape = waveform.shape
    x = waveform.reshape(-1, shape[-1]).float()
    length = x.shape[-1]
    x = torch.nn.functional.pad(x, (width, width + of))
    y = torch.nn.functional.conv1d(x[:, None], kernel, stride=of)
    y = y.transpose(1, 2).reshape(x.shape[0], -1)
    target = int(math.ceil(nf * length / of))
    return y[..., :target].reshape(*shape[:-1], target)


def _mul(samples: np.ndarray, factor: float) -> np.ndarray:
    """audioop.mul on 16-bit samples: floor(clip(sample * factor))."""
    f = np.floor(np.clip(samples.astype(np.float64) * factor, -32768.0, 32767.0))
    return f.astype(np.int16)


def pcm16_segment(audio: torch.Tensor, sample_rate: int, fade_duration: int = 50, start_silence_duration: int = 500,
                  end_silence_duration: int = 100) -> np.ndarray:
    a = audio.to(torch.float32).reshape(-1)
    a = a / max(a.abs().max(), 1e-6)
    pcm = (a.cpu().numpy() * 32767).astype("int16")
    ms = lambda d: int(d * (sample_rate / 1000.0))  # noqa: E731
    seg = np.concatenate([np.zeros(ms(start_silence_duration), np.int16), pcm, np.zeros(ms(end_silence_duration), np.int16)])
    assert fade_duration <= 100
    n = ms(fade_duration)
    lo = 10 ** (-120 / 20.0)
    out = seg.copy()
    if n > 0:
        # fade_in: from_gain -120 dB over the first n frames, one gain step per sample
        step = (1.0 - lo) / n
        for i in range(n):
            out[i] = _mul(out[i:i + 1], lo + step * i)[0]
        # fade_out: to_gain -120 dB over the last n frames
        step = (lo - 1.0) / n
        start = len(out) - n
        for i in range(n):
            out[start + i] = _mul(out[start + i:start + i + 1], 1.0 + step * i)[0]
    return out
