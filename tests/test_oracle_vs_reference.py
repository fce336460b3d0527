"""The oracle restatement must reproduce the UNMODIFIED reference ``sesameai/models.py``
(imported on the torchtune shim) bit for bit on the CPU.  Runs only where /root/reference exists."""
import torch

import csm_oracle as orc
from sesameai import synthetic as syn
from helpers import register_tiny
from torchtune.models import llama3_2 as tt


def _reference_tiny(ref):
    for name, dims in syn.named_tiny_flavors().items():
        ref.FLAVORS[name] = (lambda d=dims: tt.llama3_2(vocab_size=128_256, max_seq_len=2048, attn_dropout=0.0,
                                                          norm_eps=1e-5, rope_base=500_000, scale_factor=32, **d))


def test_oracle_equals_reference_models_py(reference_models):
    ref = reference_models
    register_tiny()
    _reference_tiny(ref)
    args = dict(backbone_flavor="tiny-bb", decoder_flavor="tiny-dec", text_vocab_size=777, audio_vocab_size=2051,
                audio_num_codebooks=32)
    rm = ref.Model(ref.ModelArgs(**args))
    om = orc.OracleCSM(orc.OracleArgs(**args))
    syn.init_random_weights(rm, 5)
    om.load_state_dict(rm.state_dict())
    assert sorted(rm.state_dict()) == sorted(om.state_dict())
    rm.to(dtype=torch.bfloat16), om.to(dtype=torch.bfloat16)
    rm.setup_caches(2), om.setup_caches(2)
    tok, msk, pos = syn.voice_prompt(2, 1, 3, 5, 4, text_vocab=777)
    noise = syn.exp_noise(32 * 5, 2, 2051)
    saved = ref._multinomial_sample_one_no_sync
    try:
        with torch.inference_mode():
            for topk, temp in ((1, 1.0), (40, 0.8)):
                calls = {"n": 0}

                def race(probs):
                    q = noise[calls["n"]]
                    calls["n"] += 1
                    return torch.argmax(probs / q, dim=-1, keepdim=True).to(dtype=torch.int)

                ref._multinomial_sample_one_no_sync = race
                want = orc.oracle_frame_loop(rm, tok, msk, pos, 5, temp, topk, stop_on_eos=False)
                got = orc.oracle_frame_loop(
                    om, tok, msk, pos, 5, temp, topk, stop_on_eos=False,
                    frame_fn=lambda i, t, m, p: om.generate_frame(t, m, p, temp, topk, noise=noise[32 * i: 32 * i + 32]))
                assert len(want) == len(got) == 5
                for a, b in zip(want, got):
                    assert a.dtype == b.dtype == torch.int32 and torch.equal(a, b)
    finally:
        ref._multinomial_sample_one_no_sync = saved


def test_oracle_sample_topk_equals_reference(reference_models):
    ref = reference_models
    logits = torch.empty(3, 2051)
    syn.hash_uniform_(logits, 1, 2, 6.0)
    logits = logits.to(torch.bfloat16)
    logits[0, 5] = logits[0].max()  # exact tie at the top survives topk=1
    q = syn.exp_noise(1, 3, 2051, 3)[0]
    saved = ref._multinomial_sample_one_no_sync
    try:
        ref._multinomial_sample_one_no_sync = lambda p: torch.argmax(p / q, dim=-1, keepdim=True).to(dtype=torch.int)
        for k, t in ((1, 1.0), (30, 0.7), (2051, 1.1)):
            assert torch.equal(ref.sample_topk(logits, k, t), orc.oracle_sample_topk(logits, k, t, q))
    finally:
        ref._multinomial_sample_one_no_sync = saved
