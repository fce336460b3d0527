"""Oracle mode (ii) of SURVEY.md 8c: the reference frame loop as tts_service.py runs it, inside
``torch.autocast("cuda", dtype=bfloat16)`` (tts_service.py:192-194), emulated on the CPU by the op dtypes of
CUDA autocast (sum / softmax / log_softmax -> fp32, linear -> bf16).  The kernels implement mode (i)
(``Generator.generate``: no autocast); these tests pin what mode (ii) changes and how far apart the modes are."""
import torch

import csm_oracle as orc
from sesameai import synthetic as syn
from helpers import build_oracle, logit_report

SPEC = dict(model_args=dict(backbone_flavor="tiny-bb", decoder_flavor="tiny-dec", text_vocab_size=1000,
                            audio_vocab_size=2051, audio_num_codebooks=32), weight_seed=5, planted=False, batch=1)


def _frame(om, mode, tok, msk, pos, noise, forced=None):
    om.autocast_cuda = mode
    om.reset_caches()
    rec = {}
    with torch.inference_mode():
        s = om.generate_frame(tok, msk, pos, 0.9, 50, noise=noise, record=rec, forced=forced)
    om.autocast_cuda = False
    return s, torch.stack(rec["logits"]).float(), torch.stack(rec["sampled"])


def test_embedding_sum_is_fp32_under_autocast():
    om, _ = build_oracle(SPEC)
    tok, msk, _ = syn.voice_prompt(1, 1, 2, 3, 1, seed=2, text_vocab=1000)
    om.autocast_cuda = True
    h2 = om.embed_frame_inputs(tok, msk)
    om.autocast_cuda = False
    h1 = om.embed_frame_inputs(tok, msk)
    assert h2.dtype == torch.float32 and h1.dtype == torch.bfloat16
    # text rows have one term: identical; audio rows sum 32 terms: the bf16 sum is the rounded fp32 sum at best
    text_rows = msk[0, :, -1]
    assert torch.equal(h2[0, text_rows].to(torch.bfloat16), h1[0, text_rows])
    assert (h2[0, ~text_rows] - h1[0, ~text_rows].float()).abs().max() > 0


def test_modes_are_one_bf16_ulp_apart_and_sample_alike():
    om, _ = build_oracle(SPEC)
    tok, msk, pos = syn.voice_prompt(1, 1, 4, 6, 3, seed=2, text_vocab=1000)
    noise = syn.exp_noise(32, 1, 2051, 3)
    s1, lg1, smp1 = _frame(om, False, tok, msk, pos, noise)
    s2, lg2, smp2 = _frame(om, True, tok, msk, pos, noise, forced=s1)
    rep = logit_report(lg2, lg1, "oracle mode ii vs mode i (tiny)")
    assert rep["cos"] >= 0.9999 and rep["max_abs"] <= 4e-2
    assert lg2.shape == lg1.shape
    # fp32 probabilities + fp32 race pick the same token as the bf16 chain almost everywhere (same logits +- 1 ulp, same noise)
    assert (smp1 == smp2).float().mean().item() >= 0.9


def test_autocast_sampling_is_fp32():
    logits = torch.randn(2, 2051).to(torch.bfloat16)
    q = syn.exp_noise(1, 2, 2051, 9)[0]
    a = orc.oracle_sample_topk(logits, 50, 0.9, q, autocast_cuda=True)
    b = orc.oracle_sample_topk(logits, 50, 0.9, q, autocast_cuda=False)
    assert a.dtype == torch.int32 and a.shape == b.shape == (2, 1)
    # greedy is mode independent
    assert torch.equal(orc.oracle_sample_topk(logits, 1, 1.0, q, True), orc.oracle_sample_topk(logits, 1, 1.0, q, False))
