"""Drop-in API conformance on the B200: ``Generator.generate`` / ``generate_stream`` and the
reference's own frame loop (tts_service.py:224-241 shape) driven against the new Model + MimiCodec,
checked against the oracle end to end (tokens bit-exact, waveform SNR >= 40 dB)."""
import math

import pytest
import torch

import csm_oracle as orc
import mimi_oracle as mo
from sesameai import synthetic as syn
from sesameai.generator import Generator, Segment
from sesameai.mimi import MimiCodec
from helpers import build_oracle, build_product

pytestmark = pytest.mark.gpu


class FakeTokenizer:
    """Deterministic stand-in for the gated Llama-3 tokenizer (not reachable offline)."""

    def encode(self, text):
        return [1] + [3 + (ord(ch) * 7919) % 990 for ch in text] + [2]


@pytest.fixture(scope="module")
def rig():
    spec = dict(model_args=dict(backbone_flavor="tiny-bb", decoder_flavor="tiny-dec", text_vocab_size=1000,
                                audio_vocab_size=2051, audio_num_codebooks=32),
                weight_seed=1234, planted=True, batch=1)
    om, _ = build_oracle(spec)
    pm, _ = build_product(spec)
    omimi = mo.OracleMimi().eval()
    syn.init_mimi_weights(omimi, 2024)
    codec = MimiCodec(max_frames=64)
    codec.load_state_dict(omimi.state_dict())
    codec.to("cuda")
    gen = Generator(pm, text_tokenizer=FakeTokenizer(), audio_tokenizer=codec)
    return om, omimi, gen


def _oracle_audio(om, omimi, gen, text, speaker, n_frames):
    tok, msk = gen._tokenize_text_segment(text, speaker)
    tok, msk = tok.cpu().unsqueeze(0), msk.cpu().unsqueeze(0)
    pos = torch.arange(tok.shape[1]).unsqueeze(0)
    with torch.inference_mode():
        frames = orc.oracle_frame_loop(om, tok, msk, pos, n_frames, 1.0, 1)
        wav = omimi.decode((torch.stack(frames).permute(1, 2, 0) % 2048))
    return torch.stack(frames), wav.squeeze(0).squeeze(0)


def test_generate_matches_oracle_end_to_end(rig):
    om, omimi, gen = rig
    # planted greedy tokens are < 2051; Mimi codebooks hold 2048 entries -> fold like the oracle side
    real_decode = gen._audio_tokenizer.decode
    gen._audio_tokenizer.decode = lambda codes: real_decode(codes % 2048)
    try:
        audio = gen.generate("hello there", speaker=0, context=[], max_audio_length_ms=12 * 80, temperature=1.0, topk=1)
    finally:
        gen._audio_tokenizer.decode = real_decode
    frames, want = _oracle_audio(om, omimi, gen, "hello there", 0, 12)
    assert audio.shape == want.shape == (12 * 1920,)
    assert audio.dtype == torch.float32 and audio.is_cuda
    got = audio.cpu()
    snr = 10 * math.log10(want.pow(2).sum().item() / (got - want).pow(2).sum().item())
    assert snr >= 40.0, snr


def test_generate_stream_chunks_and_length_guard(rig):
    _, _, gen = rig
    real_decode = gen._audio_tokenizer.decode
    gen._audio_tokenizer.decode = lambda codes: real_decode(codes % 2048)
    try:
        chunks = list(gen.generate_stream("abc", 1, [], max_audio_length_ms=23 * 80, temperature=1.0, topk=1))
        assert [c.shape[0] for c in chunks] == [19200, 19200, 3 * 1920]  # 10 + 10 + 3 frames
        whole = gen.generate("abc", 1, [], max_audio_length_ms=23 * 80, temperature=1.0, topk=1, stream=True)
        assert whole.shape[0] == 23 * 1920
    finally:
        gen._audio_tokenizer.decode = real_decode
    with pytest.raises(ValueError, match="Inputs too long"):
        gen.generate("x" * 100, 0, [], max_audio_length_ms=2000 * 80)


def test_reference_service_loop_runs_unchanged(rig):
    """The frame loop of tts_service.TTS.generate_with_context, verbatim in shape: inference_mode +
    autocast(bf16), torch.cat bookkeeping with int32 samples, torch.all EOS test."""
    _, _, gen = rig
    model = gen._model
    model.reset_caches()
    with torch.inference_mode(), torch.autocast("cuda", dtype=torch.bfloat16):
        t, m = gen._tokenize_text_segment("service", 1)
        curr_tokens, curr_mask = t.unsqueeze(0), m.unsqueeze(0)
        curr_pos = torch.arange(0, t.size(0)).unsqueeze(0).long().to("cuda")
        samples = []
        for _ in range(5):
            sample = model.generate_frame(curr_tokens, curr_mask, curr_pos, 0.9, 50)
            assert sample.dtype == torch.int32 and sample.shape == (1, 32)
            if torch.all(sample == 0):
                break
            samples.append(sample)
            curr_tokens = torch.cat([sample, torch.zeros(1, 1).long().to("cuda")], dim=1).unsqueeze(1)
            curr_mask = torch.cat([torch.ones_like(sample).bool(), torch.zeros(1, 1).bool().to("cuda")], dim=1).unsqueeze(1)
            curr_pos = curr_pos[:, -1:] + 1
        audio = gen._audio_tokenizer.decode(torch.stack(samples).permute(1, 2, 0) % 2048).squeeze(0).squeeze(0)
    assert audio.shape == (5 * 1920,) and torch.isfinite(audio).all()


def test_service_loop_values_against_the_autocast_oracle(rig):
    """VALUES of the tts_service loop: the reference runs it under autocast (oracle mode ii: fp32 backbone residual
    stream, fp32 sampling); the kernels implement the plain mode (i) whatever the autocast state.  Teacher-forced
    logits of the product, called inside torch.autocast like tts_service.py:192-241, against the mode-(ii) oracle:
    the distance is the one between the two reference modes themselves (~1 bf16 ulp), measured and bounded here."""
    from helpers import assert_logits_close, next_inputs

    spec = dict(model_args=dict(backbone_flavor="tiny-bb", decoder_flavor="tiny-dec", text_vocab_size=1000,
                                audio_vocab_size=2051, audio_num_codebooks=32), weight_seed=5, planted=False, batch=1)
    om, _ = build_oracle(spec)
    pm, _ = build_product(spec)
    tok, msk, pos = syn.voice_prompt(1, 1, 4, 6, 3, seed=2, text_vocab=1000)
    noise = syn.exp_noise(32 * 3, 1, 2051, 3)
    om.autocast_cuda = True
    om.reset_caches(), pm.reset_caches()
    tc, mc, pc = tok.cuda(), msk.cuda(), pos.cuda()
    with torch.inference_mode(), torch.autocast("cuda", dtype=torch.bfloat16):
        for f in range(3):
            rec = {}
            s = om.generate_frame(tok, msk, pos, 0.9, 50, noise=noise[32 * f: 32 * f + 32], record=rec)
            lg = torch.zeros(32, 1, 2051, dtype=torch.bfloat16, device="cuda")
            sp = pm.generate_frame(tc, mc, pc, 0.9, 50, noise=noise[32 * f: 32 * f + 32].cuda(), forced=s, logits_out=lg)
            assert torch.equal(sp.cpu(), s)
            assert_logits_close(lg.cpu(), torch.stack(rec["logits"]), f"service loop vs autocast oracle, frame {f}", worst=4e-2, ulps=2.0)
            tok, msk, pos = next_inputs(s, pos)
            tc, mc, pc = next_inputs(sp, pc)
    om.autocast_cuda = False


def test_generate_stream_equals_generate(rig):
    """Stateful streaming: the chunks of generate_stream concatenated are bit-identical to generate()'s audio."""
    _, _, gen = rig
    codec = gen._audio_tokenizer

    class Folded:  # planted tokens reach 2050; Mimi codebooks hold 2048 entries: fold like the other tests do
        sample_rate = codec.sample_rate

        def set_num_codebooks(self, n):
            pass

        def decode(self, codes):
            return codec.decode(codes % 2048)

        def streaming(self):
            st = codec.streaming()
            real = st.decode
            st.decode = lambda codes: real(codes % 2048)
            return st

    gen._audio_tokenizer = Folded()
    try:
        chunks = list(gen.generate_stream("stream me", 0, [], max_audio_length_ms=27 * 80, temperature=1.0, topk=1))
        whole = gen.generate("stream me", 0, [], max_audio_length_ms=27 * 80, temperature=1.0, topk=1)
    finally:
        gen._audio_tokenizer = codec
    assert [c.shape[0] for c in chunks] == [19200, 19200, 7 * 1920]
    assert torch.equal(torch.cat(chunks), whole)


def test_generate_with_voice_prompt_context(rig):
    """Context segments carry audio: Generator._tokenize_segment runs Mimi encode on the GPU and the
    prompt gets text frames + audio frames + the all-zero EOS frame (reference generator.py:78-109)."""
    om, omimi, gen = rig
    wav = torch.empty(24000)
    syn.hash_uniform_(wav, 12, 1, 0.4)
    seg = Segment(speaker=0, text="ctx", audio=wav)
    tok, msk = gen._tokenize_segment(seg)
    n_text = len(FakeTokenizer().encode("[0]ctx"))
    n_audio = 13 + 1  # ceil(24000 / 1920) frames + EOS frame
    assert tok.shape == (n_text + n_audio, 33) and msk.shape == tok.shape
    assert msk[:n_text, -1].all() and not msk[:n_text, :-1].any()
    assert msk[n_text:, :-1].all() and not msk[n_text:, -1].any()
    assert int(tok[-1].abs().sum()) == 0  # EOS frame
    with torch.inference_mode():
        want = omimi.encode(wav.view(1, 1, -1))[0].t()
    assert (tok[n_text:-1, :-1].cpu() == want).float().mean().item() >= 0.97
    real_decode = gen._audio_tokenizer.decode
    gen._audio_tokenizer.decode = lambda codes: real_decode(codes % 2048)
    try:
        audio = gen.generate("hi", 0, [seg], max_audio_length_ms=400, temperature=0.9, topk=50)
    finally:
        gen._audio_tokenizer.decode = real_decode
    assert audio.shape == (5 * 1920,) and torch.isfinite(audio).all()
