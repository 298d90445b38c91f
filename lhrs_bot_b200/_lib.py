"""ctypes binding of liblhrs_b200.so — the C-ABI declared in include/lhrs_b200.h.

The product path has no CPU fallback: if the shared library is missing or a symbol is absent the import
of any op fails loudly (``LhrsLibraryError``).  Nothing under ``oracle/`` is ever imported from here.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

LIB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")
LIB_PATH = os.environ.get("LHRS_LIB_PATH") or os.path.join(LIB_DIR, "liblhrs_b200.so")   # override: instrumented debug builds (tools/)


class LhrsLibraryError(RuntimeError):
    pass


class LhrsGemm(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("A", C.c_void_p), ("lda", C.c_int64), ("a_mn_major", C.c_int32),
        ("B", C.c_void_p * 3), ("num_b", C.c_int32), ("seg_rows", C.c_int32),
        ("ldb", C.c_int64), ("b_mn_major", C.c_int32),
        ("epilogue", C.c_int32), ("act", C.c_int32), ("alpha", C.c_float),
        ("bias", C.c_void_p * 3), ("residual", C.c_void_p), ("ldr", C.c_int64),
        ("D", C.c_void_p), ("ldd", C.c_int64), ("d_f32", C.c_int32),
        ("row_map", C.c_void_p),
        ("rope_cos", C.c_void_p), ("rope_sin", C.c_void_p), ("positions", C.c_void_p),
        ("rope_seq_len", C.c_int32),
        ("pre_gate", C.c_void_p), ("pre_up", C.c_void_p),
        ("A2", C.c_void_p), ("lda2", C.c_int64), ("B2", C.c_void_p * 3), ("ldb2", C.c_int64), ("ext_k", C.c_int32),
        ("b_seg_nshift", C.c_int32), ("split_k", C.c_int32), ("drop_key", C.c_uint32), ("drop_t", C.c_int32),
    ]


class LhrsAttention(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("o", C.c_void_p),
        ("lse", C.c_void_p), ("key_mask", C.c_void_p),
        ("q_bs", C.c_int64), ("q_rs", C.c_int64), ("q_hs", C.c_int64),
        ("k_bs", C.c_int64), ("k_rs", C.c_int64), ("k_hs", C.c_int64),
        ("v_bs", C.c_int64), ("v_rs", C.c_int64), ("v_hs", C.c_int64),
        ("o_bs", C.c_int64), ("o_rs", C.c_int64), ("o_hs", C.c_int64),
        ("B", C.c_int32), ("H", C.c_int32), ("Sq", C.c_int32), ("Skv", C.c_int32), ("head_dim", C.c_int32),
        ("causal", C.c_int32), ("scale", C.c_float),
        ("seq_off", C.c_void_p), ("total_rows", C.c_int64),
    ]


_PP = C.POINTER(C.c_void_p)


class LhrsVitWeights(C.Structure):
    _fields_ = [
        ("num_layers", C.c_int32), ("dim", C.c_int32), ("ffn", C.c_int32), ("heads", C.c_int32), ("patch", C.c_int32),
        ("image", C.c_int32), ("kpad", C.c_int32), ("eps", C.c_float),
        ("patch_w", C.c_void_p), ("cls", C.c_void_p), ("pos", C.c_void_p), ("pre_ln_w", C.c_void_p), ("pre_ln_b", C.c_void_p),
        ("ln1_w", _PP), ("ln1_b", _PP), ("q_w", _PP), ("q_b", _PP), ("k_w", _PP), ("k_b", _PP), ("v_w", _PP), ("v_b", _PP),
        ("o_w", _PP), ("o_b", _PP), ("ln2_w", _PP), ("ln2_b", _PP), ("fc1_w", _PP), ("fc1_b", _PP), ("fc2_w", _PP), ("fc2_b", _PP),
    ]


class LhrsPoolerWeights(C.Structure):
    _fields_ = [
        ("num_layers", C.c_int32), ("dim", C.c_int32), ("ffn", C.c_int32), ("heads", C.c_int32), ("out_dim", C.c_int32),
        ("num_groups", C.c_int32), ("stage_num", C.c_int32 * 4), ("split_part", C.c_int32 * 4), ("eps", C.c_float),
        ("query", C.c_void_p),
        ("ln1_w", _PP), ("ln1_b", _PP), ("lnkv_w", _PP), ("lnkv_b", _PP), ("in_w", _PP), ("in_b", _PP),
        ("ao_w", _PP), ("ao_b", _PP), ("ln2_w", _PP), ("ln2_b", _PP), ("fc_w", _PP), ("fc_b", _PP), ("pj_w", _PP), ("pj_b", _PP),
        ("out_w", C.c_void_p), ("out_b", C.c_void_p),
    ]


class LhrsAttentionBwd(C.Structure):
    _fields_ = [
        ("fwd", LhrsAttention), ("d_o", C.c_void_p), ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p),
        ("delta", C.c_void_p),
        ("dq_bs", C.c_int64), ("dq_rs", C.c_int64), ("dq_hs", C.c_int64),
        ("dk_bs", C.c_int64), ("dk_rs", C.c_int64), ("dk_hs", C.c_int64),
        ("dv_bs", C.c_int64), ("dv_rs", C.c_int64), ("dv_hs", C.c_int64),
        ("rope_cos", C.c_void_p), ("rope_sin", C.c_void_p),
    ]


class LhrsKvCache(C.Structure):
    _fields_ = [
        ("pool", C.c_void_p), ("block_table", C.c_void_p),
        ("layers", C.c_int32), ("heads", C.c_int32), ("head_dim", C.c_int32), ("page_size", C.c_int32),
        ("num_pages", C.c_int32), ("max_pages", C.c_int32),
    ]


class LhrsDecodeBuffers(C.Structure):
    _fields_ = [
        ("xbuf", C.c_void_p), ("qkv", C.c_void_p), ("obuf", C.c_void_p), ("act", C.c_void_p),
        ("logits", C.c_void_p), ("part_val", C.c_void_p), ("part_idx", C.c_void_p),
        ("state", C.c_void_p), ("tokens_out", C.c_void_p), ("max_tokens", C.c_int32),
        ("attn_part", C.c_void_p), ("attn_count", C.c_void_p),
    ]


class LhrsSampling(C.Structure):
    _fields_ = [
        ("do_sample", C.c_int32), ("temperature", C.c_float), ("top_k", C.c_int32), ("top_p", C.c_float),
        ("repetition_penalty", C.c_float), ("eos_token", C.c_int32), ("seed", C.c_uint64), ("seed_dev", C.c_void_p),
        ("stop_seqs", C.c_void_p), ("n_stop", C.c_int32), ("stop_len", C.c_int32), ("work", C.c_void_p),
    ]


class LhrsPeerExchange(C.Structure):
    _fields_ = [
        ("world", C.c_int32), ("rank", C.c_int32), ("grads", C.c_void_p * 16), ("params", C.c_void_p * 16),
        ("norm_slots", C.c_void_p * 16), ("slice_offset", C.c_int64), ("slice_n", C.c_int64),
        ("mc_grads", C.c_void_p), ("mc_params", C.c_void_p),
    ]


class LhrsLlamaWeights(C.Structure):
    _fields_ = [
        ("num_layers", C.c_int32), ("dim", C.c_int32), ("ffn", C.c_int32), ("heads", C.c_int32), ("vocab", C.c_int32),
        ("max_pos", C.c_int32), ("eps", C.c_float),
        ("ln1_w", _PP), ("q_w", _PP), ("k_w", _PP), ("v_w", _PP), ("o_w", _PP), ("ln2_w", _PP),
        ("gate_w", _PP), ("up_w", _PP), ("down_w", _PP),
        ("norm_w", C.c_void_p), ("lm_head", C.c_void_p), ("embed", C.c_void_p),
        ("rope_cos", C.c_void_p), ("rope_sin", C.c_void_p),
        ("lora_r", C.c_int32), ("lora_scale", C.c_float), ("lora_a", _PP), ("lora_b", _PP),
        ("lora_dropout", C.c_float), ("lora_seed", C.c_uint64),
        ("qkv_wt", _PP), ("o_wt", _PP), ("gu_wt", _PP), ("down_wt", _PP), ("lm_head_wt", C.c_void_p),
    ]


_P = C.c_void_p
_I32 = C.c_int32
_I64 = C.c_int64
_F = C.c_float

# symbol -> (restype, argtypes).  tests/test_abi.py checks this table against include/lhrs_b200.h.
SIGNATURES = {
    "lhrs_last_error": (C.c_char_p, []),
    "lhrs_version": (C.c_int, []),
    "lhrs_launch_count": (C.c_uint64, []),
    "lhrs_prof_enable": (C.c_int, [C.c_int]),
    "lhrs_prof_summary": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "lhrs_gemm_bf16": (C.c_int, [C.POINTER(LhrsGemm), _P]),
    "lhrs_cast_f32_bf16": (C.c_int, [_P, _P, _I64, _P]),
    "lhrs_attention_fwd": (C.c_int, [C.POINTER(LhrsAttention), _P]),
    "lhrs_rmsnorm_fwd": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _F, _P]),
    "lhrs_layernorm_fwd": (C.c_int, [_P, _I64, _P, _P, _P, _I64, _P, _P, _I64, _I32, _F, _P]),
    "lhrs_vit_im2col": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I32, _P]),
    "lhrs_vit_embed_ln": (C.c_int, [_P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _F, _P]),
    "lhrs_splice_scan": (C.c_int, [_P, _I32, _I32, _I32, _P, _P]),
    "lhrs_splice_fill": (C.c_int, [_P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P]),
    "lhrs_splice_bwd": (C.c_int, [_P, _P, _P, _I64, _I32, _P]),
    "lhrs_ce_fwd": (C.c_int, [_P, _I64, _P, _I32, _I32, _I32, _P, _P, _P, _P]),
    "lhrs_ce_bwd": (C.c_int, [_P, _I64, _P, _I32, _I32, _I32, _P, _P, _F, _P, _P, _P]),
    "lhrs_vit_workspace_bytes": (C.c_size_t, [C.POINTER(LhrsVitWeights), _I32]),
    "lhrs_vit_fwd": (C.c_int, [C.POINTER(LhrsVitWeights), _P, _I32, C.POINTER(C.c_int32), _I32, _P, _P, C.c_size_t, _P]),
    "lhrs_pooler_workspace_bytes": (C.c_size_t, [C.POINTER(LhrsPoolerWeights), _I32]),
    "lhrs_pooler_stash_bytes": (C.c_size_t, [C.POINTER(LhrsPoolerWeights), _I32]),
    "lhrs_pooler_fwd": (C.c_int, [C.POINTER(LhrsPoolerWeights), _P, _I32, _P, _I64, _P, _P, _P, C.c_size_t, _P]),
    "lhrs_llama_workspace_bytes": (C.c_size_t, [C.POINTER(LhrsLlamaWeights), _I32, _I32]),
    "lhrs_llama_stash_bytes": (C.c_size_t, [C.POINTER(LhrsLlamaWeights), _I32, _I32]),
    "lhrs_llama_fwd": (C.c_int, [C.POINTER(LhrsLlamaWeights), _P, _I32, _I32, _P, _P, _P, C.POINTER(LhrsKvCache), _P, C.c_size_t, _P]),
    "lhrs_llama_fwd_ragged": (C.c_int, [C.POINTER(LhrsLlamaWeights), _P, _I32, _I32, _I64, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "lhrs_lm_head": (C.c_int, [C.POINTER(LhrsLlamaWeights), _P, _I64, _P, _P]),
    "lhrs_llama_first_token": (C.c_int, [C.POINTER(LhrsLlamaWeights), _P, _I32, C.POINTER(LhrsDecodeBuffers), _I32, _P]),
    "lhrs_llama_decode_step": (C.c_int, [C.POINTER(LhrsLlamaWeights), C.POINTER(LhrsKvCache), C.POINTER(LhrsDecodeBuffers), _I32, _I32, _P]),
    "lhrs_decode_commit_token": (C.c_int, [C.POINTER(LhrsLlamaWeights), C.POINTER(LhrsDecodeBuffers), _I32, _I32, _P]),
    "lhrs_clip_preprocess_workspace_bytes": (C.c_size_t, [_I32, _I32, _I32, _I32]),
    "lhrs_clip_preprocess": (C.c_int, [_P, _I32, _I32, _I32, _I32, C.POINTER(C.c_float), C.POINTER(C.c_float), _P, _I32, _P, _P, C.c_size_t, _P]),
    "lhrs_sample_logits": (C.c_int, [_P, _I32, _P, _I32, C.POINTER(LhrsSampling), C.c_uint64, _P, _P, _P]),
    "lhrs_llama_first_token_sampled": (C.c_int, [C.POINTER(LhrsLlamaWeights), _P, _I32, C.POINTER(LhrsDecodeBuffers), C.POINTER(LhrsSampling), _P]),
    "lhrs_llama_decode_step_sampled": (C.c_int, [C.POINTER(LhrsLlamaWeights), C.POINTER(LhrsKvCache), C.POINTER(LhrsDecodeBuffers),
                                                 C.POINTER(LhrsSampling), _I32, _P]),
    "lhrs_attention_bwd": (C.c_int, [C.POINTER(LhrsAttentionBwd), _P]),
    "lhrs_attention_bwd_grouped": (C.c_int, [C.POINTER(LhrsAttentionBwd), _I32, _P]),
    "lhrs_attention_fwd_grouped": (C.c_int, [C.POINTER(LhrsAttention), _I32, _P]),
    "lhrs_attention_bwd_scratch_floats": (C.c_int64, [_I32, _I32, _I32, _I32]),
    "lhrs_lora_dropout_mask": (C.c_int, [_P, _I64, _I64, _I32, C.c_uint64, _I32, _F, _P, _I64, _P]),
    "lhrs_lora_dx_dropout": (C.c_int, [_P, _I64, _I64, _I32, _P, _I64, _PP, _I32, C.c_uint64, _I32, _F, _P]),
    "lhrs_lora_panel_dropout": (C.c_int, [_P, _I64, _I64, _I32, _P, _I64, _I32, _F, C.c_uint64, _I32, _F, _P, _I64, _P]),
    "lhrs_lora_rowreduce_dropout": (C.c_int, [_P, _I64, _I64, _I32, _P, _I64, _I32, _P, _I64, C.c_uint64, _I32, _F, _P, C.c_size_t, _P]),
    "lhrs_rmsnorm_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _I64, _I32, _P]),
    "lhrs_layernorm_bwd_scratch_bytes": (C.c_size_t, [_I32]),
    "lhrs_layernorm_bwd": (C.c_int, [_P, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _P, _I64, _I32, _P]),
    "lhrs_colsum_scratch_bytes": (C.c_size_t, [_I32]),
    "lhrs_colsum": (C.c_int, [_P, _I64, _I64, _I32, _P, _I32, _P, _P]),
    "lhrs_swiglu_bwd": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _P]),
    "lhrs_gelu_bwd": (C.c_int, [_P, _P, _I64, _P]),
    "lhrs_rope_bwd": (C.c_int, [_P, _I64, _I64, _I32, _P, _P, _P, _I32, _P]),
    "lhrs_lora_panel": (C.c_int, [_P, _I64, _I64, _I32, _PP, _I32, _I32, _I64, _I32, _F, _P, _I64, _P]),
    "lhrs_lora_rowreduce_scratch_bytes": (C.c_size_t, [_I64, _I32, _I32]),
    "lhrs_lora_rowreduce": (C.c_int, [_P, _I64, _I64, _I32, _P, _I64, _I32, _I32, _I32, _PP, _I64, _F, _P, C.c_size_t, _P]),
    "lhrs_llama_bwd_workspace_bytes": (C.c_size_t, [C.POINTER(LhrsLlamaWeights), _I32, _I32]),
    "lhrs_llama_bwd": (C.c_int, [C.POINTER(LhrsLlamaWeights), _PP, _PP, _P, _I32, _I32, _P, _P, _P, _P, C.c_size_t, _P]),
    "lhrs_llama_bwd_ragged": (C.c_int, [C.POINTER(LhrsLlamaWeights), _PP, _PP, _P, _I32, _I32, _I64, _P, _P, _P, _P, C.c_size_t, _P]),
    "lhrs_lm_head_bwd": (C.c_int, [C.POINTER(LhrsLlamaWeights), _P, _I64, _P, _P]),
    "lhrs_pooler_bwd_workspace_bytes": (C.c_size_t, [C.POINTER(LhrsPoolerWeights), _I32]),
    "lhrs_pooler_bwd": (C.c_int, [C.POINTER(LhrsPoolerWeights), C.POINTER(LhrsPoolerWeights), _P, _I64, _I32, _P, _P, _P, C.c_size_t, _P]),
    "lhrs_grad_sumsq": (C.c_int, [_P, _I64, _P, _P, _P]),
    "lhrs_adamw_step": (C.c_int, [_P, _P, _P, _P, _P, _P, _I64, _F, _F, _F, _F, _F, _I32, _P, _F, _F, _P]),
    "lhrs_p2p_reduce_slice": (C.c_int, [C.POINTER(LhrsPeerExchange), _P, _P, _P]),
    "lhrs_p2p_adamw_slice": (C.c_int, [C.POINTER(LhrsPeerExchange), _P, _P, _P, _P, _P, _F, _F, _F, _F, _F, _I32, _F, _F, _P]),
    "lhrs_p2p_adan_slice": (C.c_int, [C.POINTER(LhrsPeerExchange), _P, _P, _P, _P, _P, _P, _P, _F, _F, _F, _F, _F, _F, _I32, _I32, _F, _F, _P]),
    "lhrs_adan_step": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _F, _F, _F, _F, _F, _F, _I32, _I32, _P, _F, _F, _P]),
}

_lock = threading.Lock()
_lib = None


def load() -> C.CDLL:
    """Load the shared library once; raise (never fall back) if it is missing or incomplete."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise LhrsLibraryError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                f"(or `make -C lhrs_bot_b200/csrc`). There is no CPU fallback for this path.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as e:
                raise LhrsLibraryError(f"{LIB_PATH} does not export {name}") from e
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().lhrs_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")
