"""GPU post-processing (sesameai.postprocess) vs the oracle: resampling within fp32 rounding of the conv1d
restatement, the PCM segment bit-exact (integer work)."""
import numpy as np
import pytest
import torch

import post_oracle as po
from sesameai import postprocess as pp
from sesameai import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,src,dst", [(48000, 24000, 44100), (88200, 44100, 24000), (1, 24000, 44100), (1921, 24000, 16000),
                                       (24000 * 20, 24000, 44100)])
def test_resample_matches_oracle(n, src, dst):
    x = torch.empty(n)
    syn.hash_uniform_(x, 3, n % 1013, 0.8)
    want = po.resample(x, src, dst)
    got = pp.resample(x.cuda(), src, dst).cpu()
    assert got.shape == want.shape
    assert (got - want).abs().max().item() <= 5e-6 * max(1.0, want.abs().max().item()), (got - want).abs().max().item()


def test_resample_batch_and_identity():
    x = torch.empty(3, 5000)
    syn.hash_uniform_(x, 5, 1, 0.5)
    got = pp.resample(x.cuda(), 24000, 44100).cpu()
    for b in range(3):
        assert torch.equal(got[b], pp.resample(x[b].cuda(), 24000, 44100).cpu())
    xc = x.cuda()
    assert pp.resample(xc, 24000, 24000) is xc


@pytest.mark.parametrize("n,fade,s0,s1", [(24000 * 3, 50, 500, 100), (5000, 50, 0, 0), (1, 0, 1, 1), (24000, 0, 500, 100),
                                          (2000, 100, 20, 20)])
def test_pcm16_segment_bit_exact(n, fade, s0, s1):
    a = torch.empty(n)
    syn.hash_uniform_(a, 7, n % 911, 0.37)
    want = po.pcm16_segment(a, 24000, fade, s0, s1)
    got = pp.pcm16_segment(a.cuda(), 24000, fade, s0, s1).cpu().numpy()
    assert got.shape == want.shape
    assert np.array_equal(got, want), int((got != want).sum())


def test_pcm16_segment_of_silence():
    got = pp.pcm16_segment(torch.zeros(2000, device="cuda"), 24000, 50, 10, 10)
    assert int(got.abs().max()) == 0
