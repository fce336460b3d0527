"""a16 (``load_csm_1b`` / ``Model.from_pretrained``): the product ``Model`` round-trips through the
PyTorchModelHubMixin checkpoint format the reference loads ``sesame/csm-1b`` with (reference
generator.py:338) -- from a local directory, since the hub is not reachable offline -- with the reference's
state-dict keys."""
import dataclasses
import os

import pytest
import torch

from helpers import register_tiny
from sesameai import synthetic as syn
from sesameai.models import Model, ModelArgs

ARGS = dict(backbone_flavor="tiny-bb", decoder_flavor="tiny-dec", text_vocab_size=1000, audio_vocab_size=2051,
            audio_num_codebooks=32)


def test_save_and_from_pretrained_local_dir(tmp_path):
    pytest.importorskip("huggingface_hub")
    pytest.importorskip("safetensors")
    register_tiny()
    m = Model(ModelArgs(**ARGS))
    syn.init_random_weights(m, 77)
    m.save_pretrained(str(tmp_path))
    assert os.path.exists(tmp_path / "config.json") and os.path.exists(tmp_path / "model.safetensors")
    m2 = Model.from_pretrained(str(tmp_path))
    assert dataclasses.asdict(m2.config) == ARGS if dataclasses.is_dataclass(m2.config) else True
    sd, sd2 = m.state_dict(), m2.state_dict()
    assert sorted(sd) == sorted(sd2)
    for k in sd:
        assert torch.equal(sd[k], sd2[k]), k
    # the reference's key names (sesameai/models.py:113-118 + torchtune module tree)
    for k in ("text_embeddings.weight", "audio_embeddings.weight", "projection.weight", "codebook0_head.weight", "audio_head",
              "backbone.layers.0.attn.q_proj.weight", "backbone.layers.1.mlp.w2.weight", "decoder.layers.0.sa_norm.scale",
              "decoder.norm.scale"):
        assert k in sd, k


@pytest.mark.gpu
def test_load_csm_1b_from_local_checkpoint(tmp_path):
    """The loader end to end on the GPU: checkpoint -> bf16 model on the device -> Generator with caches -> a frame."""
    from sesameai.generator import load_csm_1b
    from sesameai.mimi import MimiCodec

    register_tiny()
    m = Model(ModelArgs(**ARGS))
    syn.init_random_weights(m, 78)
    m.save_pretrained(str(tmp_path))

    class Tok:
        def encode(self, text):
            return [1] + [3 + (ord(c) * 31) % 900 for c in text] + [2]

    codec = MimiCodec(max_frames=16)
    syn.init_mimi_weights(codec, 2024)
    codec.to("cuda")
    gen = load_csm_1b("cuda", model_path=str(tmp_path), text_tokenizer=Tok(), audio_tokenizer=codec)
    assert next(gen._model.parameters()).dtype == torch.bfloat16 and gen.sample_rate == 24000
    audio = gen.generate("hi", 0, [], max_audio_length_ms=3 * 80, temperature=0.9, topk=50)
    assert audio.shape == (3 * 1920,) and torch.isfinite(audio).all()
