"""CPU checks of the oracle against the committed golden vectors (outputs of the unmodified
reference ``sesameai/models.py`` run on the shim, tests/golden/make_golden.py)."""
import torch

import csm_oracle as orc
from sesameai import synthetic as syn
from helpers import build_oracle, gold_inputs, load_golden


def test_sample_topk_known_answers():
    for case in load_golden("sample_topk_cases.pt"):
        got = orc.oracle_sample_topk(case["logits"], case["topk"], case["temperature"], case["noise"])
        assert torch.equal(got.view(-1).to(torch.int32), case["token"])


@torch.inference_mode()
def test_tiny_greedy_tokens_and_planted_chain():
    gold = load_golden("tiny_greedy.pt")
    om, perms = build_oracle(gold)
    tok, msk, pos, noise = gold_inputs(gold)
    F = gold["frames"].shape[0]
    frames = orc.oracle_frame_loop(
        om, tok, msk, pos, F, 1.0, 1,
        frame_fn=lambda i, t, m, p: om.generate_frame(t, m, p, 1.0, 1, noise=noise[32 * i: 32 * i + 32]))
    got = torch.stack(frames)
    assert torch.equal(got, gold["frames"])
    assert gold["min_margin_ulps"] >= 8
    # the planted structure makes the greedy sequence analytic: c0 = pi0^-1(c31 of the previous
    # frame), ci = pi_i^-1(c_{i-1}); for the first frame "c31" is the last text token mod V
    V = 2051
    prev = (tok[:, -1, -1] % V).tolist()
    for f in range(F):
        for b in range(got.shape[1]):
            want = syn.planted_next_frame(perms, prev[b])
            assert torch.equal(got[f, b], want), (f, b)
            prev[b] = int(got[f, b, -1])


@torch.inference_mode()
def test_tiny_teacher_forced_logits():
    gold = load_golden("tiny_teacher.pt")
    om, _ = build_oracle(gold)
    tok, msk, pos, noise = gold_inputs(gold)
    F = gold["frames"].shape[0]
    om.reset_caches()
    for f in range(F):
        rec = {}
        s = om.generate_frame(tok, msk, pos, gold["temperature"], gold["topk"], noise=noise[32 * f: 32 * f + 32],
                              forced=gold["frames"][f], record=rec)
        lg = torch.stack(rec["logits"]).float()  # [32, B, V]
        want = gold["logits"][f].float()
        assert (lg - want).abs().max() <= 2e-2
        cos = torch.nn.functional.cosine_similarity(lg.flatten(), want.flatten(), dim=0)
        assert cos >= 0.999
        assert torch.equal(torch.stack(rec["sampled"]).squeeze(-1).t().to(torch.int32), gold["frames"][f])
        from helpers import next_inputs

        tok, msk, pos = next_inputs(s, pos)
