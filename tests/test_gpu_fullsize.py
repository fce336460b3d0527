"""Full-size (CSM-1B) parity at the BASELINE batch shapes, against goldens of the unmodified reference models.py
(tests/golden/make_golden.py): config 3's 1568-frame voice prompt (tensor-core prefill + batched decode at that
context) at B = 2 and tiled to B = 32, a B = 64 decode step (tcgen05 decode GEMMs), and Mimi decode of 60 s
utterances (T = 750) against the oracle.  Streams of a batch are independent, so a golden generated for B
streams pins any batch that repeats those streams."""
import math
import os

import pytest
import torch

from sesameai import _native
from sesameai import synthetic as syn
from helpers import GOLDEN, assert_logits_close, build_product, load_golden, next_inputs

pytestmark = pytest.mark.gpu
COS_MIN = 0.999


def _teacher_forced_tiled(pm, gold, tok, msk, pos, reps, prefill=0):
    """Run gold['frames'] frames under teacher forcing on a batch that repeats the golden's streams ``reps``
    times; yields (frame index, logits [32, B, V] fp32 cpu, sampled tokens [B, 32])."""
    Bg = gold["batch"]
    B = Bg * reps
    F = gold["frames"].shape[0]
    noise = syn.exp_noise(32 * F, Bg, 2051, gold["noise_seed"]).cuda()
    t, m, p = tok.repeat(reps, 1, 1).cuda(), msk.repeat(reps, 1, 1).cuda(), pos.repeat(reps, 1).cuda()
    pm.reset_caches()
    for f in range(F):
        lg = torch.zeros(32, B, 2051, dtype=torch.bfloat16, device="cuda")
        smp = torch.zeros(B, 32, dtype=torch.int32, device="cuda")
        forced = gold["frames"][f].repeat(reps, 1)
        s = pm.generate_frame(t, m, p, gold["temperature"], gold["topk"], noise=noise[32 * f: 32 * f + 32].repeat(1, reps, 1),
                              forced=forced, logits_out=lg, sampled_out=smp, prefill=prefill if f == 0 else 0)
        assert torch.equal(s.cpu(), forced)
        yield f, lg.float().cpu(), smp.cpu()
        t, m, p = next_inputs(s, p)
    torch.cuda.synchronize()
    pm.check_device_error()


def _check_against_gold(gold, f, lg, smp, reps, label):
    Bg = gold["batch"]
    want = gold["logits"][f].float()  # [32, Bg, V]
    agree = 0
    for r in range(reps):
        got = lg[:, r * Bg:(r + 1) * Bg]
        assert_logits_close(got, want, f"{label} frame {f} copy {r}")
        for cb in range(32):
            c = torch.nn.functional.cosine_similarity(got[cb].flatten(), want[cb].flatten(), dim=0).item()
            assert c >= COS_MIN, (f, r, cb, c)
        agree += int((smp[r * Bg:(r + 1) * Bg] == gold["frames"][f]).sum())
    # sampled ids equal wherever the logits agree closely enough (random weights: a few near-ties may flip)
    assert agree >= 0.9 * reps * Bg * 32, agree


@pytest.fixture(scope="module")
def voice_gold():
    path = os.path.join(GOLDEN, "csm1b_voice1568.pt")
    if not os.path.exists(path):
        pytest.skip("csm1b_voice1568.pt has not been generated (tests/golden/make_golden.py voice)")
    return load_golden("csm1b_voice1568.pt")


@pytest.mark.parametrize("reps", [1, 16])
def test_config3_voice_prompt_prefill_and_decode(voice_gold, reps):
    """1568-frame prompt per stream: tensor-core prefill (TMA + tcgen05 GEMMs, tiled attention), then one decode
    step at that context -- B = 2 (skinny GEMM decode) and B = 32 (the BASELINE config 3 batch)."""
    gold = voice_gold
    pm, _ = build_product(gold, batch=gold["batch"] * reps)
    tok, msk, pos = syn.voice_prompt(gold["batch"], **gold["prompt"])
    assert tok.shape[1] == 1568
    for f, lg, smp in _teacher_forced_tiled(pm, gold, tok, msk, pos, reps, prefill=_native.PREFILL_TENSOR):
        _check_against_gold(gold, f, lg, smp, reps, f"config3 B={gold['batch'] * reps}")


def test_decode_B64_on_the_tcgen05_path():
    """64 streams (the reference's 24-frame teacher prompt repeated): prompt rows and every decode GEMM run on
    the TMA + tcgen05 kernels (128 rows at depth step 1); three teacher-forced frames against the golden."""
    gold = load_golden("csm1b_teacher.pt")
    reps = 64 // gold["batch"]
    pm, _ = build_product(gold, batch=gold["batch"] * reps)
    tok, msk, pos = syn.text_prompt(gold["batch"], gold["prompt_frames"], gold["input_seed"])
    for f, lg, smp in _teacher_forced_tiled(pm, gold, tok, msk, pos, reps):
        _check_against_gold(gold, f, lg, smp, reps, "B=64")


def test_mimi_decode_60s_batch4_vs_oracle():
    """BASELINE config 4's utterance length (T = 750 frames = 60 s, 1500 transformer positions: six times the
    250-key window) at batch 4, against the oracle decoder (SNR >= 60 dB; north_star gate 40 dB)."""
    import mimi_oracle as mo
    from sesameai.mimi import MimiCodec

    om = mo.OracleMimi().eval()
    syn.init_mimi_weights(om, 2024)
    pc = MimiCodec(max_frames=760)
    pc.load_state_dict(om.state_dict())
    pc.to("cuda")
    B, T = 4, 750
    codes = syn.hash_ints(B * 32 * T, 41, T, 2048).view(B, 32, T)
    got = pc.decode(codes.cuda()).cpu()
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.inference_mode():
        want = om.decode(codes)
    assert got.shape == want.shape == (B, 1, 1920 * T)
    for b in range(B):
        err = (got[b].double() - want[b].double()).pow(2).sum().item()
        snr = 10 * math.log10(want[b].double().pow(2).sum().item() / max(err, 1e-300))
        print(f"[mimi 60 s] utterance {b}: SNR {snr:.1f} dB")
        assert snr >= 60.0, (b, snr)
