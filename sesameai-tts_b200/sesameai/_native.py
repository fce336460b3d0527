"""ctypes binding of libcsm_b200.so (include/csm_b200.h).

There is deliberately NO fallback here: if the shared library has not been built, or a call
fails, an exception is raised.  Build with ``python sesameai-tts_b200/build.py`` (or
``__graft_entry__.build()``)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("CSM_B200_LIB") or os.path.join(_PKG_ROOT, "lib", "libcsm_b200.so")  # env: kernel-variant experiments

MIMI_W_CODEBOOK0, MIMI_W_RVQ_FIRST_PROJ, MIMI_W_RVQ_REST_PROJ, MIMI_W_UPSAMPLE = 0, 64, 65, 66
MIMI_W_LAYER0, MIMI_W_CONV0, MIMI_W_STAGE0, MIMI_W_FINAL, MIMI_W_COUNT = 67, 147, 149, 173, 286
PREFILL_AUTO, PREFILL_SMALL_ROW, PREFILL_TENSOR = 0, 1, 2
PATH_AUTO, PATH_DIRECT, PATH_GRAPH, PATH_MEGA = 0, 1, 2, 3
CSM_OK, CSM_ERR_ARG, CSM_ERR_CUDA, CSM_ERR_STATE, CSM_ERR_OVERFLOW, CSM_ERR_WORKSPACE = 0, -1, -2, -3, -4, -5


class StackConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("layers", "dim", "heads", "kv_heads", "ff")]


class Config(C.Structure):
    _fields_ = [
        ("backbone", StackConfig),
        ("decoder", StackConfig),
        ("text_vocab", C.c_int32),
        ("audio_vocab", C.c_int32),
        ("codebooks", C.c_int32),
        ("max_seq_len", C.c_int32),
        ("norm_eps", C.c_float),
    ]


class PhaseInfo(C.Structure):
    """csm_phase_info (include/csm_b200.h): one row of the megakernel's phase table, host-only debug view."""
    _fields_ = ([(n, C.c_int32) for n in ("type", "epi", "nb", "K", "rows", "R", "G", "rot", "ldx", "ldo", "split_row",
                                           "attn_prologue", "has_qkv_table")]
                + [("x_src", C.c_int32 * 2), ("resid_src", C.c_int32 * 2), ("q_src", C.c_int32), ("logits_src", C.c_int32)]
                + [(n, C.c_uint64) for n in ("t_x", "t_out", "t_out2", "t_q", "t_kv", "t_logits", "t_next")]
                + [(n, C.c_int32) for n in ("kv_sync", "done_src", "pos0", "pos_mode")])


class LayerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("q_proj", "k_proj", "v_proj", "output_proj", "w1", "w2", "w3", "sa_norm", "mlp_norm")]


class Weights(C.Structure):
    _fields_ = [
        ("text_embeddings", C.c_void_p),
        ("audio_embeddings", C.c_void_p),
        ("projection", C.c_void_p),
        ("codebook0_head", C.c_void_p),
        ("audio_head", C.c_void_p),
        ("backbone_norm", C.c_void_p),
        ("decoder_norm", C.c_void_p),
        ("backbone_rope", C.c_void_p),
        ("decoder_rope", C.c_void_p),
        ("backbone_rope_len", C.c_int32),
        ("decoder_rope_len", C.c_int32),
        ("backbone_layers", C.POINTER(LayerWeights)),
        ("decoder_layers", C.POINTER(LayerWeights)),
    ]


class FrameOpts(C.Structure):
    _fields_ = [
        ("noise", C.c_void_p),
        ("seed", C.c_uint64),
        ("offset", C.c_uint64),
        ("forced", C.c_void_p),
        ("logits_out", C.c_void_p),
        ("sampled_out", C.c_void_p),
        ("path", C.c_int32),
        ("prefill", C.c_int32),
        ("lanes", C.POINTER(C.c_int32)),
    ]


class NativeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libcsm_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg


_lib: Optional[C.CDLL] = None

# symbol -> (restype, argtypes); every symbol include/csm_b200.h declares
PROTOTYPES = {
    "csm_abi_version": (C.c_int32, []),
    "csm_last_error": (C.c_char_p, []),
    "csm_launch_count": (C.c_uint64, []),
    "csm_debug_set_pdl": (None, [C.c_int32]),
    "csm_workspace_bytes": (C.c_size_t, [C.POINTER(Config), C.c_int32]),
    "csm_create": (C.c_int32, [C.POINTER(Config), C.POINTER(Weights), C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p,
                               C.POINTER(C.c_void_p)]),
    "csm_destroy": (None, [C.c_void_p]),
    "csm_reset_caches": (C.c_int32, [C.c_void_p]),
    "csm_cache_len": (C.c_int32, [C.c_void_p]),
    "csm_check_error": (C.c_int32, [C.c_void_p, C.c_int32]),
    "csm_lane_reset": (C.c_int32, [C.c_void_p, C.c_int32]),
    "csm_lane_len": (C.c_int32, [C.c_void_p, C.c_int32]),
    "csm_generate_frame": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float,
                                       C.c_int32, C.POINTER(FrameOpts), C.c_void_p, C.c_void_p]),
    "mimi_workspace_bytes": (C.c_size_t, [C.c_int32]),
    "mimi_create": (C.c_int32, [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p,
                                C.POINTER(C.c_void_p)]),
    "mimi_decode": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "mimi_encode": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "mimi_k_rvq_encode": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "mimi_destroy": (None, [C.c_void_p]),
    "mimi_stream_state_bytes": (C.c_size_t, []),
    "mimi_stream_create": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_void_p)]),
    "mimi_stream_reset": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "mimi_decode_stream": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "mimi_stream_destroy": (None, [C.c_void_p]),
    "csm_post_resample_len": (C.c_int64, [C.c_int64, C.c_int32, C.c_int32]),
    "csm_post_resample_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "csm_post_resample": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "csm_post_pcm16_segment": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                           C.c_void_p, C.c_void_p]),
    "csm_debug_set_trace": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "csm_debug_phase_table": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]),
    "csm_k_sample_topk": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_void_p,
                                      C.c_void_p]),
    "csm_k_embed_frames": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_int32, C.c_void_p, C.c_void_p]),
    "csm_k_linear": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "csm_k_gemm_tc": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                  C.c_void_p]),
    "csm_k_gemm_tc_splitk": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "csm_k_attn_prefill": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_int32, C.c_void_p, C.c_void_p]),
    "csm_k_rmsnorm": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p]),
}


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built and this package has no "
                "fallback path. Run `python sesameai-tts_b200/build.py`."
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.csm_abi_version() != 1:
            raise RuntimeError("libcsm_b200.so ABI version mismatch; rebuild")
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != CSM_OK:
        raise NativeError(rc, lib().csm_last_error().decode("utf-8", "replace"))
