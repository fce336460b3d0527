"""Continuous batching on the B200: streams with different prompt lengths join and leave KV-cache lanes at
different steps; every request must get exactly the frames of the oracle's single-stream loop
(reference generator.py:283-294) -- planted greedy weights, so the tokens are well-posed in bf16."""
import pytest
import torch

import csm_oracle as orc
from sesameai import synthetic as syn
from sesameai.serving import ContinuousBatcher, LaneGroups, Request
from helpers import build_oracle, build_product, next_inputs

pytestmark = pytest.mark.gpu

SPEC = dict(model_args=dict(backbone_flavor="tiny-bb", decoder_flavor="tiny-dec", text_vocab_size=1000,
                            audio_vocab_size=2051, audio_num_codebooks=32),
            weight_seed=1234, planted=True, batch=1)


def _requests(n):
    reqs = []
    for r in range(n):
        tok, msk, _ = syn.text_prompt(1, 3 + (r * 2) % 5, 100 + r, 1000)
        reqs.append(Request(r, tok[0], msk[0], 3 + (r * 3) % 4))
    return reqs


def _oracle_frames(om, req):
    tok, msk = req.tokens.unsqueeze(0), req.mask.unsqueeze(0)
    pos = torch.arange(tok.shape[1]).unsqueeze(0)
    with torch.inference_mode():
        out = orc.oracle_frame_loop(om, tok, msk, pos, req.max_frames, 1.0, 1)
    return torch.cat(out, dim=0) if out else torch.zeros(0, 32, dtype=torch.int32)


@pytest.mark.parametrize("groups", [1, 2])
def test_streams_join_and_leave_lanes(groups):
    om, _ = build_oracle(SPEC)
    pm, _ = build_product(SPEC, batch=1)
    reqs = _requests(7)
    want = {r.rid: _oracle_frames(om, r) for r in reqs}
    if groups == 1:
        b = ContinuousBatcher(pm, 3, 1.0, 1)
        got = b.run(reqs)
        assert b.steps > 0 and b.row_steps > b.steps  # several lanes advanced per decode call
    else:
        got = LaneGroups(pm, 2, 2, 1.0, 1).run(reqs)
    torch.cuda.synchronize()
    for r in reqs:
        assert torch.equal(got[r.rid], want[r.rid]), r.rid


def test_lane_api_on_the_model():
    """Two lanes at different lengths advance in one call; a lane other than 0 works at batch 1 (megakernel)."""
    om, _ = build_oracle(SPEC)
    pm, _ = build_product(SPEC, batch=3)
    reqs = _requests(2)
    want = {r.rid: _oracle_frames(om, r) for r in reqs}
    pm.reset_caches()
    state = {}
    for r, lane in ((reqs[0], 2), (reqs[1], 0)):
        tok, msk = r.tokens.cuda().unsqueeze(0), r.mask.cuda().unsqueeze(0)
        pos = torch.arange(tok.shape[1], device="cuda").unsqueeze(0)
        s = pm.generate_frame(tok, msk, pos, 1.0, 1, lanes=[lane])
        assert torch.equal(s.cpu()[0], want[r.rid][0])
        state[lane] = (r, s)
    assert pm.lane_len(2) == reqs[0].tokens.shape[0] and pm.lane_len(0) == reqs[1].tokens.shape[0] and pm.lane_len(1) == 0
    lanes = [2, 0]
    s = torch.cat([state[l][1] for l in lanes])
    pos = torch.tensor([[pm.lane_len(l)] for l in lanes], device="cuda")
    tok, msk, _ = next_inputs(s, pos)
    out = pm.generate_frame(tok, msk, pos, 1.0, 1, lanes=lanes).cpu()
    for i, l in enumerate(lanes):
        assert torch.equal(out[i], want[state[l][0].rid][1])
    with pytest.raises(Exception):
        pm.generate_frame(tok, msk, pos, 1.0, 1, lanes=[1, 1])


@pytest.mark.parametrize("batch", [3, 20, 40])
def test_dependent_launch_matches_plain_launches(batch):
    """The row-batched path launches its kernels with the programmatic-dependent-launch attribute (kernel n+1
    starts while kernel n drains, csrc/common.cuh).  Sampled frames (temperature 0.9, top-k 50: any stale
    activation changes the tokens) must be identical to plain stream-ordered launches, through the captured
    graph and through direct launches: skinny kernels (3, 20 rows) and tcgen05 GEMMs (40 rows)."""
    from sesameai import _native
    spec = dict(SPEC, planted=False)
    runs = {}
    try:
        for pdl in (1, 0):
            _native.lib().csm_debug_set_pdl(pdl)
            pm, _ = build_product(spec, batch=batch)  # a fresh context: its graphs are captured under this setting
            for direct in (False, True):
                torch.manual_seed(5)
                pm.reset_caches()
                tok, msk, pos = syn.text_prompt(batch, 9, 11, 1000, device="cuda")
                frames = []
                for _ in range(24):
                    s = pm.generate_frame(tok, msk, pos, 0.9, 50, no_graph=direct)
                    frames.append(s)
                    tok, msk, pos = next_inputs(s, pos)
                runs[(pdl, direct)] = torch.stack(frames).cpu()
                pm.check_device_error()
    finally:
        _native.lib().csm_debug_set_pdl(1)
    assert len(torch.unique(runs[(0, False)])) > 100  # really sampled
    for key, got in runs.items():
        assert torch.equal(got, runs[(0, True)]), key
