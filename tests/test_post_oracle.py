"""The post-processing oracle on the CPU: audioop pin for the fade arithmetic, properties for the resampler
(torchaudio is not installed in this image: that part is parity-unpinned, see oracle/post_oracle.py)."""
import math
import warnings

import numpy as np
import pytest
import torch

import post_oracle as po


def test_mul_matches_cpython_audioop():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        audioop = pytest.importorskip("audioop")
    rng = np.random.default_rng(0)
    s = rng.integers(-32768, 32768, 4000).astype(np.int16)
    for factor in (1e-6, 0.123456, 0.5, 0.999999, 1.0, 1.7):
        want = np.frombuffer(audioop.mul(s.tobytes(), 2, factor), dtype=np.int16)
        assert np.array_equal(po._mul(s, factor), want), factor


def test_resample_properties():
    sr = 24000
    t = torch.arange(sr // 2) / sr
    tone = torch.sin(2 * math.pi * 1000.0 * t)
    up = po.resample(tone, 24000, 44100)
    assert up.shape[-1] == math.ceil(44100 * tone.shape[-1] / 24000)
    # the tone keeps its frequency and amplitude (away from the edges)
    tt = torch.arange(up.shape[-1]) / 44100
    ref = torch.sin(2 * math.pi * 1000.0 * tt)
    assert (up[500:-500] - ref[500:-500]).abs().max() < 2e-3
    back = po.resample(up, 44100, 24000)
    assert back.shape[-1] == tone.shape[-1]
    assert (back[300:-300] - tone[300:-300]).abs().max() < 5e-3
    assert po.resample(tone, 24000, 24000) is tone
    dc = po.resample(torch.ones(4000), 24000, 44100)
    assert (dc[200:-200] - 1).abs().max() < 1e-3


def test_segment_layout():
    a = torch.linspace(-0.5, 0.25, 2400)
    seg = po.pcm16_segment(a, 24000, fade_duration=50, start_silence_duration=500, end_silence_duration=100)
    assert seg.dtype == np.int16 and len(seg) == 12000 + 2400 + 2400
    assert (seg[:12000] == 0).all() and (seg[-2400:] == 0).all()
    assert seg[12000] == -32767 and seg[12000 + 2399] == int(0.5 * 32767)  # peak-normalised, truncated
    # fades reach into the audio when the silences are shorter than the fade
    seg2 = po.pcm16_segment(a, 24000, fade_duration=50, start_silence_duration=0, end_silence_duration=0)
    assert seg2[0] == math.floor(-32767 * 1e-6) and abs(int(seg2[1199])) < abs(int(seg[12000 + 1199]))
