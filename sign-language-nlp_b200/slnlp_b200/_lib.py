"""ctypes binding of libslnlp_b200.so (the C ABI in include/slnlp_b200.h).

There is no CPU fallback: if the shared library is missing, importing this module
raises; if a call fails, ``check`` raises RuntimeError with the library's message.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_uint32, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libslnlp_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build the CUDA extension first "
        "(python sign-language-nlp_b200/build.py or __graft_entry__.build()). "
        "slnlp_b200 has no CPU fallback.")

lib = ctypes.CDLL(LIB_PATH)

P, I, L, F, U32 = c_void_p, c_int, c_int64, c_float, c_uint32

# name -> argtypes (restype int unless listed in _RESTYPES); mirrors include/slnlp_b200.h
SIGNATURES = {
    "slnlp_abi_version": [],
    "slnlp_last_error_string": [],
    "slnlp_device_sm_count": [],
    "slnlp_stream_create": [],
    "slnlp_stream_create_priority": [I],
    "slnlp_stream_destroy": [P],
    "slnlp_launch_count": [],
    "slnlp_max_active_clusters": [I, I],
    "slnlp_debug_persist_config": [I, I, P],
    "slnlp_embed_gather_fwd": [P, P, P, I, I, I, P, P, P, I, F, P, P],
    "slnlp_embed_gather_bwd": [P, P, P, I, I, I, P, P, P, I, F, L, P],
    "slnlp_gemm_f32": [I, I, I, I, I, P, I, P, I, P, I, P, F, P, L, P],
    "slnlp_gemm_workspace_floats": [],
    "slnlp_gemm_tf32": [I, I, I, I, I, P, I, P, I, P, I, P, F, P, L, P],
    "slnlp_gemm_tf32x3": [I, I, I, I, I, P, I, P, I, P, I, P, F, P, L, P],
    "slnlp_gemm_bf16_supported": [I, I, I, I, I],
    "slnlp_gemm_bf16": [I, I, I, I, I, P, L, P, L, P, I, P, F, P],
    "slnlp_cast_bf16": [P, L, P, L, L, L, I, P],
    "slnlp_colsum_f32": [P, I, I, I, P, F, P],
    "slnlp_colsum_bf16": [P, I, I, L, P, F, P],
    "slnlp_dropout_bf16": [P, P, L, F, P, U32, P],
    "slnlp_dropout_bf16_masked": [P, P, P, L, F, P, U32, P],
    "slnlp_rnn_layer_fwd": [I, I, I, I, I, I, P, P, P, P, P, P, P, P, P, P],
    "slnlp_rnn_layer_bwd": [I, I, I, I, I, I, P, P, P, P, P, P, P, P, P, P, P, P, P, P],
    "slnlp_rnn_extras_supported": [I, I, I, I, I],
    "slnlp_rnn_bf16_step_supported": [I, I, I, I, I],
    "slnlp_rnn_layer_fwd_bf16": [I, I, I, I, I, P, P, P, P, P, P, P, P, P],
    "slnlp_rnn_layer_bwd_bf16": [I, I, I, I, I, P, P, P, P, P, P, P, P, P, P, I, P, F, I, P],
    "slnlp_rnn_layer_fwd_bf16_ex": [I, I, I, I, I, P, P, P, P, P, P, P, P, P, P],
    "slnlp_rnn_bf16_pair_supported": [I, I, I, I, I],
    "slnlp_rnn_layer_fwd_ex": [I, I, I, I, I, I, P, P, P, P, P, P, P, P, P, P, P],
    "slnlp_rnn_layer_bwd_ex": [I, I, I, I, I, I, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P],
    "slnlp_dec_cell_fwd": [I, I, I, I, P, P, P, P, P, P, P, P, P, P, P, F, P, U32, P],
    "slnlp_dec_cell_bwd_supported": [I, I, I, I],
    "slnlp_dec_cell_bwd": [I, I, I, I, P, P, P, P, P, P, P, P, P, P, P, F, P, U32, P],
    "slnlp_dec_head_supported": [I, I, I, I, I, I],
    "slnlp_dec_head_fwd": [I, I, I, I, I, P, P, P, P, P, P, P, P, L, P, P, P, P, P, P, P],
    "slnlp_dec_head_bwd": [I, I, I, I, I, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P],
    "slnlp_pad_fill": [P, P, I, I, I, F, P],
    "slnlp_pad_fill_copy": [P, P, P, I, I, I, F, P],
    "slnlp_concat_dirs": [P, P, I, I, I, I, P],
    "slnlp_tanh_fwd": [P, L, P],
    "slnlp_tanh_bwd": [P, P, L, P],
    "slnlp_relu_fwd": [P, L, P],
    "slnlp_relu_bwd": [P, P, L, P],
    "slnlp_relu_dropout_fwd": [P, L, F, P, U32, P],
    "slnlp_relu_dropout_bwd": [P, P, L, F, P],
    "slnlp_dropout": [P, P, L, F, P, U32, P],
    "slnlp_rng_advance": [P, P],
    "slnlp_dropout_mask": [P, L, F, P, U32, P],
    "slnlp_dec_input_fwd": [P, P, P, I, I, I, P],
    "slnlp_dec_input_bwd": [P, P, P, I, I, I, P],
    "slnlp_axpy": [P, P, F, L, P],
    "slnlp_attn_step_fwd": [P, P, P, P, P, L, I, I, I, I, P, P, P],
    "slnlp_attn_step_bwd": [P, P, P, P, P, P, I, I, I, I, P, P, P, P, P],
    "slnlp_log_softmax_fwd": [P, P, I, I, P],
    "slnlp_log_softmax_bwd": [P, P, P, I, I, I, P],
    "slnlp_ce_on_logp": [P, P, L, I, I, P, P, I, P, P],
    "slnlp_logsoftmax_ce_fused": [P, P, L, I, I, P, P, P, I, P, P],
    "slnlp_ce_reduce": [P, I, P, P],
    "slnlp_sumsq_partials": [],
    "slnlp_gradnorm": [P, L, P, P, P],
    "slnlp_sgd_momentum_clip": [P, P, P, L, P, P, F, P],
    "slnlp_sgd_momentum_clip_zero": [P, P, P, L, P, P, F, P],
    "slnlp_mha_fwd": [P, I, P, I, P, I, P, I, P, I, I, I, I, I, I, P, L, F, P, U32, P],
    "slnlp_mha_bwd": [P, I, P, I, P, I, P, P, I, P, P, P, P, P, I, I, I, I, I, I, P, L, F, P, U32, P],
    "slnlp_mha_tf32_fwd": [P, I, P, I, P, I, P, I, P, I, I, I, I, I, I, P, L, F, P, U32, P],
    "slnlp_mha_tf32_bwd": [P, I, P, I, P, I, P, P, I, P, P, P, P, P, I, I, I, I, I, I, P, L, F, P, U32, P],
    "slnlp_add_layernorm_fwd": [P, P, P, P, P, P, P, I, I, F, P],
    "slnlp_ln_bwd_blocks": [I],
    "slnlp_layernorm_bwd": [P, P, P, P, P, P, P, P, I, I, I, P],
}
_RESTYPES = {"slnlp_last_error_string": c_char_p, "slnlp_launch_count": c_int64, "slnlp_stream_create": c_void_p, "slnlp_stream_create_priority": c_void_p,
             "slnlp_gemm_workspace_floats": c_int64}

for _name, _args in SIGNATURES.items():
    try:
        _fn = getattr(lib, _name)
    except AttributeError:
        continue  # declared for a later build stage; tests check header <-> exports
    _fn.argtypes = _args
    _fn.restype = _RESTYPES.get(_name, c_int)


class RnnExtras(ctypes.Structure):
    """slnlp_rnn_extras (include/slnlp_b200.h)."""
    _fields_ = [("hfinal_cat", c_int), ("out_drop", c_void_p), ("p_drop", c_float), ("rng", c_void_p),
                ("site", c_uint32), ("dout_dropped", c_int), ("mask", c_void_p)]


def last_error():
    s = lib.slnlp_last_error_string()
    return s.decode() if s else ""


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError(f"slnlp_b200 {what} failed: {last_error()}")
