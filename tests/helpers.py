"""Shared builders for the parity tests."""
import os

import torch

import csm_oracle as orc
from sesameai import synthetic as syn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def register_tiny():
    orc.ARCH.update(syn.named_tiny_flavors())
    from sesameai import models as prod

    for k, v in syn.named_tiny_flavors().items():
        prod.register_flavor(k, **v)


def build_oracle(gold, batch=None):
    """Oracle model with the weights a golden file was generated from (bf16, caches on)."""
    register_tiny()
    m = orc.OracleCSM(orc.OracleArgs(**gold["model_args"]))
    syn.init_random_weights(m, gold["weight_seed"], residual_out_scale=0.1 if gold["planted"] else 1.0)
    perms = syn.plant_greedy_structure(m, gold["weight_seed"]) if gold["planted"] else None
    m.to(dtype=torch.bfloat16)
    m.setup_caches(batch or gold["batch"])
    return m, perms


def build_product(gold, device="cuda", batch=None):
    """Product model (CUDA, bf16) with the same weights; generated on the device by the
    device-independent hash fill, so they are bit-identical to the oracle's."""
    register_tiny()
    from sesameai.models import Model, ModelArgs

    m = Model(ModelArgs(**gold["model_args"]))
    m.to(device=device)
    syn.init_random_weights(m, gold["weight_seed"], residual_out_scale=0.1 if gold["planted"] else 1.0)
    perms = syn.plant_greedy_structure(m, gold["weight_seed"]) if gold["planted"] else None
    m.to(dtype=torch.bfloat16)
    m.setup_caches(batch or gold["batch"])
    return m, perms


def gold_inputs(gold, device="cpu"):
    tv = gold["model_args"]["text_vocab_size"]
    tok, msk, pos = syn.text_prompt(gold["batch"], gold["prompt_frames"], gold["input_seed"], tv, device=device)
    n_frames = gold["frames"].shape[0]
    noise = syn.exp_noise(32 * n_frames, gold["batch"], 2051, gold["noise_seed"], device=device)
    return tok, msk, pos, noise


def next_inputs(sample, pos):
    """Frame-loop bookkeeping of Generator.generate (reference generator.py:290-294)."""
    B = sample.shape[0]
    dev = sample.device
    tok = torch.cat([sample.long(), torch.zeros(B, 1, dtype=torch.long, device=dev)], dim=1).unsqueeze(1)
    msk = torch.cat([torch.ones_like(sample).bool(), torch.zeros(B, 1, dtype=torch.bool, device=dev)], dim=1).unsqueeze(1)
    return tok, msk, pos[:, -1:] + 1
