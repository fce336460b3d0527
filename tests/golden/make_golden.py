"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
``/root/reference/sesameai/models.py`` (imported on top of ``oracle/shim``: torchtune is not
vendored by the reference nor installed here) on the CPU in bf16, exactly as
``load_csm_1b`` casts the model (reference ``generator.py:343``).

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py tiny        # seconds
    python tests/golden/make_golden.py csm1b       # minutes (1.55 B parameters on the CPU)

Inputs are regenerated from seeds by ``sesameai.synthetic`` (device-independent hash fill), so
only seeds + outputs are stored.  The Exp(1) race noise of ``sample_topk`` is injected by
replacing the module attribute ``_multinomial_sample_one_no_sync`` at run time with a version
that reads the shared noise tensor (the reference file itself is not touched).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "sesameai-tts_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "oracle", "shim")):
    sys.path.insert(0, p)

import csm_oracle as orc  # noqa: E402
from sesameai import synthetic as syn  # noqa: E402
from torchtune.models import llama3_2 as tt  # noqa: E402  (shim)

WEIGHT_SEED, INPUT_SEED, NOISE_SEED = 1234, 4321, 777


def load_reference_models():
    spec = importlib.util.spec_from_file_location("_reference_sesameai_models", "/root/reference/sesameai/models.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    orc.ARCH.update(syn.named_tiny_flavors())
    for name, dims in syn.named_tiny_flavors().items():
        mod.FLAVORS[name] = (lambda d=dims: tt.llama3_2(vocab_size=128_256, max_seq_len=2048, attn_dropout=0.0,
                                                          norm_eps=1e-5, rope_base=500_000, scale_factor=32, **d))
    return mod


def margins_ulps(logits: torch.Tensor) -> torch.Tensor:
    """top-2 margin of each row in bf16 ulps of the winner."""
    top = torch.topk(logits.float(), 2, dim=-1).values
    ulp = torch.exp2(torch.floor(torch.log2(top[..., 0].abs().clamp_min(1e-30))) - 7)
    return (top[..., 0] - top[..., 1]) / ulp


class NoisePatch:
    """Feeds sample_topk's exponential race from a shared tensor, call by call."""

    def __init__(self, mod, noise):
        self.mod, self.noise, self.i = mod, noise, 0
        self.logits = []

    def __enter__(self):
        self.saved = self.mod._multinomial_sample_one_no_sync

        def race(probs):
            q = self.noise[self.i]
            self.i += 1
            return torch.argmax(probs / q, dim=-1, keepdim=True).to(dtype=torch.int)

        self.mod._multinomial_sample_one_no_sync = race
        return self

    def __exit__(self, *a):
        self.mod._multinomial_sample_one_no_sync = self.saved


def build_reference(mod, flavor, text_vocab, planted, batch):
    args = dict(backbone_flavor=flavor[0], decoder_flavor=flavor[1], text_vocab_size=text_vocab,
                audio_vocab_size=2051, audio_num_codebooks=32)
    m = mod.Model(mod.ModelArgs(**args))
    syn.init_random_weights(m, WEIGHT_SEED, residual_out_scale=0.1 if planted else 1.0)
    if planted:
        syn.plant_greedy_structure(m, WEIGHT_SEED)
    m.to(dtype=torch.bfloat16)
    m.setup_caches(batch)
    return m, args


@torch.inference_mode()
def greedy_case(mod, flavor, text_vocab, batch, prompt_frames, n_frames, out_path):
    t0 = time.time()
    m, args = build_reference(mod, flavor, text_vocab, planted=True, batch=batch)
    tok, msk, pos = syn.text_prompt(batch, prompt_frames, INPUT_SEED, text_vocab)
    noise = syn.exp_noise(32 * n_frames, batch, 2051, NOISE_SEED)
    # record logits through a hook on sample_topk's input (module attribute swap, file untouched)
    rec = []
    orig = mod.sample_topk

    def spy(logits, topk, temperature):
        rec.append(logits.detach().clone())
        return orig(logits, topk, temperature)

    mod.sample_topk = spy
    try:
        with NoisePatch(mod, noise):
            frames = orc.oracle_frame_loop(m, tok, msk, pos, n_frames, 1.0, 1)
    finally:
        mod.sample_topk = orig
    frames = torch.stack(frames)  # [F, B, 32]
    mg = margins_ulps(torch.stack(rec))  # [F*32, B]
    assert frames.shape[0] == n_frames, "hit EOS"
    assert mg.min() >= 8, f"planted margin too small: {mg.min()}"
    torch.save(dict(model_args=args, weight_seed=WEIGHT_SEED, input_seed=INPUT_SEED, noise_seed=NOISE_SEED,
                    planted=True, batch=batch, prompt_frames=prompt_frames, temperature=1.0, topk=1,
                    frames=frames.to(torch.int32), min_margin_ulps=float(mg.min())), out_path)
    print(f"{out_path}: {n_frames} frames, min margin {mg.min():.1f} ulps, {time.time() - t0:.1f}s")


@torch.inference_mode()
def teacher_case(mod, flavor, text_vocab, batch, prompt_frames, n_frames, temperature, topk, out_path):
    """Plain random weights; the reference free-runs with shared noise; we store its tokens and
    the logits at every sampling call, to be compared under teacher forcing."""
    t0 = time.time()
    m, args = build_reference(mod, flavor, text_vocab, planted=False, batch=batch)
    tok, msk, pos = syn.text_prompt(batch, prompt_frames, INPUT_SEED, text_vocab)
    noise = syn.exp_noise(32 * n_frames, batch, 2051, NOISE_SEED)
    rec = []
    orig = mod.sample_topk

    def spy(logits, k, t):
        rec.append(logits.detach().clone())
        return orig(logits, k, t)

    mod.sample_topk = spy
    try:
        with NoisePatch(mod, noise):
            frames = orc.oracle_frame_loop(m, tok, msk, pos, n_frames, temperature, topk, stop_on_eos=False)
    finally:
        mod.sample_topk = orig
    frames = torch.stack(frames)
    logits = torch.stack(rec).view(n_frames, 32, batch, 2051)
    # "truth": the same bf16 parameter values evaluated in fp32 arithmetic, teacher-forced with the
    # reference's tokens (oracle restatement, validated against the reference in bf16 above)
    om = orc.OracleCSM(orc.OracleArgs(**args))
    om.load_state_dict({k: v.float() for k, v in m.state_dict().items() if not k.endswith("causal_mask")})
    om.setup_caches(batch)
    om.reset_caches()
    t32, m32, p32 = tok, msk, pos
    truth = []
    for f in range(n_frames):
        r = {}
        s = om.generate_frame(t32, m32, p32, temperature, topk, noise=noise[32 * f: 32 * f + 32].float(),
                              forced=frames[f], record=r)
        truth.append(torch.stack(r["logits"]))
        t32 = torch.cat([s.long(), torch.zeros(batch, 1).long()], dim=1).unsqueeze(1)
        m32 = torch.cat([torch.ones_like(s).bool(), torch.zeros(batch, 1).bool()], dim=1).unsqueeze(1)
        p32 = p32[:, -1:] + 1
    truth = torch.stack(truth)
    ref_rms = (logits.float() - truth).pow(2).mean().sqrt().item()
    torch.save(dict(model_args=args, weight_seed=WEIGHT_SEED, input_seed=INPUT_SEED, noise_seed=NOISE_SEED,
                    planted=False, batch=batch, prompt_frames=prompt_frames, temperature=temperature, topk=topk,
                    frames=frames.to(torch.int32), logits=logits, logits_fp32=truth.to(torch.float32),
                    ref_rms_vs_fp32=ref_rms), out_path)
    print(f"  bf16 reference vs fp32 arithmetic: rms {ref_rms:.5f}, max {(logits.float() - truth).abs().max():.5f}")
    print(f"{out_path}: {n_frames} frames, {time.time() - t0:.1f}s")


@torch.inference_mode()
def voice_prompt_case(mod, batch, n_frames, temperature, topk, out_path):
    """BASELINE config 3's prompt on the full-size model: ``batch`` streams x 1568 frames (4 x (64 text + 320
    audio frames ending in the all-zero EOS frame) + 32 text), the reference free-runs ``n_frames`` frames with
    shared noise; tokens + logits are stored for teacher-forced comparison.  Streams are independent, so the
    GPU tests also tile these ``batch`` prompts to larger batches (B = 32)."""
    t0 = time.time()
    m, args = build_reference(mod, ("llama-1B", "llama-100M"), 128_256, planted=False, batch=batch)
    tok, msk, pos = syn.voice_prompt(batch, 4, 64, 320, 32, seed=3)
    assert tok.shape[1] == 1568
    noise = syn.exp_noise(32 * n_frames, batch, 2051, NOISE_SEED)
    rec = []
    orig = mod.sample_topk

    def spy(logits, k, t):
        rec.append(logits.detach().clone())
        return orig(logits, k, t)

    mod.sample_topk = spy
    try:
        with NoisePatch(mod, noise):
            frames = orc.oracle_frame_loop(m, tok, msk, pos, n_frames, temperature, topk, stop_on_eos=False)
    finally:
        mod.sample_topk = orig
    frames = torch.stack(frames)
    logits = torch.stack(rec).view(n_frames, 32, batch, 2051)
    torch.save(dict(model_args=args, weight_seed=WEIGHT_SEED, noise_seed=NOISE_SEED, planted=False, batch=batch,
                    prompt=dict(segments=4, text_frames=64, audio_frames=320, tail_text_frames=32, seed=3),
                    temperature=temperature, topk=topk, frames=frames.to(torch.int32), logits=logits), out_path)
    print(f"{out_path}: {n_frames} frames after a 1568-frame prompt, batch {batch}, {time.time() - t0:.1f}s")


@torch.inference_mode()
def sample_cases(mod, out_path):
    """Known-answer vectors for sample_topk alone (reference models.py:77-87)."""
    cases = []
    g = 0
    for scale in (0.5, 4.0, 20.0):
        for temperature, topk in ((1.0, 1), (0.7, 30), (0.8, 40), (0.9, 50), (1.3, 2051), (0.9, 2)):
            logits = torch.empty(4, 2051)
            syn.hash_uniform_(logits, 99, g, scale * 3 ** 0.5)
            logits = logits.to(torch.bfloat16)
            if g % 3 == 0:  # plant exact ties at the top and at the k-th value
                logits[0, 7] = logits[0].max()
                logits[1, 100:104] = torch.topk(logits[1].float(), min(topk, 2051)).values[-1].to(torch.bfloat16)
            q = syn.exp_noise(1, 4, 2051, 1000 + g)[0]
            with NoisePatch(mod, q.unsqueeze(0)):
                tok = mod.sample_topk(logits, topk, temperature)
            cases.append(dict(logits=logits, noise=q, temperature=temperature, topk=topk, token=tok.view(-1).to(torch.int32)))
            g += 1
    torch.save(cases, out_path)
    print(f"{out_path}: {len(cases)} cases")


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    mod = load_reference_models()
    torch.manual_seed(0)
    if what == "teacher":
        teacher_case(mod, ("llama-1B", "llama-100M"), 128_256, 1, 24, 3, 0.9, 50, os.path.join(HERE, "csm1b_teacher.pt"))
    elif what == "tiny":
        sample_cases(mod, os.path.join(HERE, "sample_topk_cases.pt"))
        greedy_case(mod, ("tiny-bb", "tiny-dec"), 1000, 2, 7, 12, os.path.join(HERE, "tiny_greedy.pt"))
        teacher_case(mod, ("tiny-bb", "tiny-dec"), 1000, 2, 7, 4, 0.8, 40, os.path.join(HERE, "tiny_teacher.pt"))
    elif what == "csm1b":
        greedy_case(mod, ("llama-1B", "llama-100M"), 128_256, 1, 24, 64, os.path.join(HERE, "csm1b_greedy.pt"))
        teacher_case(mod, ("llama-1B", "llama-100M"), 128_256, 1, 24, 3, 0.9, 50, os.path.join(HERE, "csm1b_teacher.pt"))
    elif what == "voice":
        voice_prompt_case(mod, 2, 2, 0.9, 50, os.path.join(HERE, "csm1b_voice1568.pt"))
    elif what == "mimi":
        return
    else:
        raise SystemExit("usage: make_golden.py [tiny|csm1b|teacher|voice|mimi]")


if __name__ == "__main__":
    main()


@torch.inference_mode()
def mimi_case(out_path):
    """Mimi decode golden: the waveform the independent ``transformers`` MimiModel port produces
    from the seeded synthetic codec weights (moshi itself is not installable offline)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import mimi_oracle as mo
    from test_mimi_oracle_pin_hf import map_to_hf, tf_mimi
    from transformers import MimiConfig

    om = mo.OracleMimi().eval()
    syn.init_mimi_weights(om, 2024)
    cfg = MimiConfig()
    cfg._attn_implementation = "eager"
    hf = tf_mimi.MimiModel(cfg).eval()
    map_to_hf(om, hf)
    B, T = 2, 20
    codes = syn.hash_ints(B * 32 * T, 7, T, 2048).view(B, 32, T)
    wav = hf.decode(codes)[0]
    torch.save(dict(weight_seed=2024, B=B, T=T, code_seed=7, wav=wav.float()), out_path)
    print(f"{out_path}: {tuple(wav.shape)}")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "mimi":
    mimi_case(os.path.join(HERE, "mimi_decode.pt"))
