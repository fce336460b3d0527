"""Pin the oracle's fp32 mathematics against an INDEPENDENT implementation of the same
published model: the ``transformers`` CSM port (HF:csm/modeling_csm.py).  The two differ in
RoPE convention (rotate-half with permuted q/k rows vs interleaved pairs), cache layout and
module structure, so agreement pins layer order, GQA grouping, Llama-3 RoPE scaling, RMSNorm
eps, SwiGLU wiring, codebook offsets and depth-decoder positions."""
import pytest
import torch

import csm_oracle as orc
from sesameai import synthetic as syn

tf_csm = pytest.importorskip("transformers.models.csm.modeling_csm")
from transformers.models.csm.configuration_csm import CsmConfig, CsmDepthDecoderConfig  # noqa: E402

ROPE = dict(rope_type="llama3", factor=32.0, low_freq_factor=1.0, high_freq_factor=4.0,
            original_max_position_embeddings=8192, rope_theta=500000.0)


def _to_rotate_half(w: torch.Tensor, n_heads: int) -> torch.Tensor:
    out, inp = w.shape
    return w.view(n_heads, out // n_heads // 2, 2, inp).transpose(1, 2).reshape(out, inp)


def _load_stack(hf_layers, hf_norm, ora_stack, n_heads, n_kv):
    sd = ora_stack.state_dict()
    for i, layer in enumerate(hf_layers):
        g = lambda k: sd[f"layers.{i}.{k}"]  # noqa: E731
        layer.self_attn.q_proj.weight.data.copy_(_to_rotate_half(g("attn.q_proj.weight"), n_heads))
        layer.self_attn.k_proj.weight.data.copy_(_to_rotate_half(g("attn.k_proj.weight"), n_kv))
        layer.self_attn.v_proj.weight.data.copy_(g("attn.v_proj.weight"))
        layer.self_attn.o_proj.weight.data.copy_(g("attn.output_proj.weight"))
        layer.mlp.gate_proj.weight.data.copy_(g("mlp.w1.weight"))
        layer.mlp.up_proj.weight.data.copy_(g("mlp.w3.weight"))
        layer.mlp.down_proj.weight.data.copy_(g("mlp.w2.weight"))
        layer.input_layernorm.weight.data.copy_(g("sa_norm.scale"))
        layer.post_attention_layernorm.weight.data.copy_(g("mlp_norm.scale"))
    hf_norm.weight.data.copy_(sd["norm.scale"])


@pytest.fixture(scope="module")
def tiny_pair():
    tiny = syn.named_tiny_flavors()
    orc.ARCH.update(tiny)
    args = orc.OracleArgs("tiny-bb", "tiny-dec", text_vocab_size=512, audio_vocab_size=2051, audio_num_codebooks=32)
    om = orc.OracleCSM(args)
    syn.init_random_weights(om, 99)
    om.setup_caches(2)
    bb, dec = tiny["tiny-bb"], tiny["tiny-dec"]
    dcfg = CsmDepthDecoderConfig(
        num_codebooks=32, backbone_hidden_size=bb["embed_dim"], vocab_size=2051, hidden_size=dec["embed_dim"],
        intermediate_size=dec["intermediate_dim"], num_hidden_layers=dec["num_layers"],
        num_attention_heads=dec["num_heads"], num_key_value_heads=dec["num_kv_heads"], rms_norm_eps=1e-5,
        max_position_embeddings=33, rope_parameters=dict(ROPE))
    cfg = CsmConfig(
        num_codebooks=32, vocab_size=2051, text_vocab_size=512, hidden_size=bb["embed_dim"],
        intermediate_size=bb["intermediate_dim"], num_hidden_layers=bb["num_layers"],
        num_attention_heads=bb["num_heads"], num_key_value_heads=bb["num_kv_heads"], rms_norm_eps=1e-5,
        max_position_embeddings=2048, rope_parameters=dict(ROPE), depth_decoder_config=dcfg)
    cfg._attn_implementation = "eager"
    dcfg._attn_implementation = "eager"
    hf_bb = tf_csm.CsmBackboneModel(cfg).eval()
    hf_dd = tf_csm.CsmDepthDecoderModel(dcfg).eval()
    _load_stack(hf_bb.layers, hf_bb.norm, om.backbone, bb["num_heads"], bb["num_kv_heads"])
    _load_stack(hf_dd.layers, hf_dd.norm, om.decoder, dec["num_heads"], dec["num_kv_heads"])
    hf_dd.inputs_embeds_projector.weight.data.copy_(om.projection.weight.data)
    return om, hf_bb, hf_dd


@torch.inference_mode()
def test_backbone_matches_hf_port(tiny_pair):
    om, hf_bb, _ = tiny_pair
    tok, msk, pos = syn.voice_prompt(2, 1, 5, 9, 4, text_vocab=512)
    om.reset_caches()
    emb = om.embed_frame_inputs(tok, msk)
    want = hf_bb(inputs_embeds=emb, use_cache=False).last_hidden_state
    got = om.backbone(emb, input_pos=pos, mask=om.backbone_causal_mask[pos, :])
    assert torch.allclose(got, want, atol=2e-4, rtol=1e-4), (got - want).abs().max()
    # incremental decode through the oracle's KV cache equals one-shot causal evaluation
    om.reset_caches()
    S = tok.shape[1]
    a = om.backbone(emb[:, : S - 3], input_pos=pos[:, : S - 3], mask=om.backbone_causal_mask[pos[:, : S - 3], :])
    outs = [a]
    for t in range(S - 3, S):
        outs.append(om.backbone(emb[:, t : t + 1], input_pos=pos[:, t : t + 1], mask=om.backbone_causal_mask[pos[:, t : t + 1], :]))
    assert torch.allclose(torch.cat(outs, 1), want, atol=2e-4, rtol=1e-4)


@torch.inference_mode()
def test_depth_decoder_matches_hf_port(tiny_pair):
    om, _, hf_dd = tiny_pair
    B, C, V = 2, 32, 2051
    codes = syn.hash_ints(B * C, 5, 3, V).view(B, C)
    last_h = torch.empty(B, om.projection.in_features)
    syn.hash_uniform_(last_h, 5, 4, 1.0)
    # oracle: incremental, exactly as generate_frame drives it (models.py:165-182)
    om.decoder.reset_caches()
    cur = torch.cat([last_h.unsqueeze(1), om.embed_audio(0, codes[:, :1])], dim=1)
    pos = torch.arange(2).unsqueeze(0).repeat(B, 1)
    got = []
    for i in range(1, C):
        dh = om.decoder(om.projection(cur), input_pos=pos, mask=om.decoder_causal_mask[pos, :])
        got.append(dh[:, -1])
        cur = om.embed_audio(i, codes[:, i : i + 1])
        pos = pos[:, -1:] + 1
    got = torch.stack(got, 1)  # positions 1..31
    # HF: one causal pass over [last_h, emb(c0), ..., emb(c30)]
    embs = torch.stack([last_h] + [om.embed_audio(i, codes[:, i]) for i in range(C - 1)], 1)
    want = hf_dd(inputs_embeds=embs, use_cache=False).last_hidden_state[:, 1:]
    assert torch.allclose(got, want, atol=2e-4, rtol=1e-4), (got - want).abs().max()


@torch.inference_mode()
def test_hf_embedding_offsets(tiny_pair):
    om, hf_bb, _ = tiny_pair
    hf_bb.embed_tokens.embed_audio_tokens.weight.data.copy_(om.audio_embeddings.weight.data)
    codes = syn.hash_ints(2 * 3 * 32, 8, 1, 2051).view(2, 3, 32)
    tok = torch.cat([codes, torch.zeros(2, 3, 1, dtype=torch.long)], -1)
    msk = torch.ones_like(tok, dtype=torch.bool)
    msk[..., -1] = False
    assert torch.allclose(om.embed_frame_inputs(tok, msk), hf_bb.embed_tokens(codes), atol=1e-5)
