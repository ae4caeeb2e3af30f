"""ctypes binding of the C ABI in include/cpt_b200.h (cpt_b200/lib/libcpt_b200.so).

There is no fallback: if the shared library is missing or fails to load, importing the engine raises.
Build it with `python -c "import __graft_entry__ as g; g.build()"` or `make -C cpt_b200/csrc`.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libcpt_b200.so")
ABI_VERSION = 3
K_COUNT = 22  # CPT_K_COUNT


class CptError(RuntimeError):
    """Raised for every non-zero status of the C ABI (a RuntimeError, so the reference's per-step
    `except RuntimeError: continue` at Oscar/oscar/fewshot/refcoco_cpt.py:244-253 still works)."""


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "hidden_size", "num_hidden_layers", "num_attention_heads", "intermediate_size", "vocab_size",
        "max_position_embeddings", "type_vocab_size", "img_feature_dim", "use_img_layernorm",
        "num_contrast_classes")] + [("layer_norm_eps", C.c_float), ("img_layer_norm_eps", C.c_float),
                                    ("dtype", C.c_int32)]


_fp = C.c_void_p
LAYER_FIELDS = ("q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "ao_w", "ao_b", "ao_ln_g", "ao_ln_b", "i_w", "i_b",
                "o_w", "o_b", "o_ln_g", "o_ln_b")
GLOBAL_FIELDS = ("word_emb", "pos_emb", "type_emb", "emb_ln_g", "emb_ln_b", "img_w", "img_b", "img_ln_g",
                 "img_ln_b", "pooler_w", "pooler_b", "mlm_dense_w", "mlm_dense_b", "mlm_ln_g", "mlm_ln_b",
                 "mlm_bias", "nsp_w", "nsp_b")


class LayerWeights(C.Structure):
    _fields_ = [(n, _fp) for n in LAYER_FIELDS]


class Weights(C.Structure):
    _fields_ = [(n, _fp) for n in GLOBAL_FIELDS] + [("layers", C.POINTER(LayerWeights))]


GRAD_GLOBAL_FIELDS = ("word_emb", "pos_emb", "type_emb", "emb_ln_g", "emb_ln_b", "img_w", "img_b", "img_ln_g",
                      "img_ln_b", "mlm_dense_w", "mlm_dense_b", "mlm_ln_g", "mlm_ln_b", "mlm_bias", "pooler_w",
                      "pooler_b", "nsp_w", "nsp_b")


class LayerGrads(C.Structure):
    _fields_ = [(n, _fp) for n in LAYER_FIELDS]


class Grads(C.Structure):
    _fields_ = [(n, _fp) for n in GRAD_GLOBAL_FIELDS] + [("layers", C.POINTER(LayerGrads))]


class Dropout(C.Structure):
    _fields_ = [("p_hidden", C.c_float), ("p_attn", C.c_float), ("seed", C.c_uint64), ("seed_dev", C.c_void_p)]


class ChainStage(C.Structure):
    """cpt_chain_stage: one stage of the dataflow chain kernel (kind 0 GEMM, 1 LayerNorm)."""
    _fields_ = [(n, C.c_int32) for n in ("kind", "M", "N", "K", "gelu", "out_fp32", "ksplit", "dep_stage")] + [
        ("A", C.c_void_p), ("lda", C.c_int64), ("W", C.c_void_p), ("ldw", C.c_int64), ("bias", C.c_void_p),
        ("out", C.c_void_p), ("ldo", C.c_int64), ("ln_in", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("eps", C.c_float), ("out32", C.c_void_p), ("out16", C.c_void_p), ("ln", C.c_int32), ("resid", C.c_void_p),
        ("ldr", C.c_int64), ("part", C.c_void_p), ("rpart", C.c_void_p), ("apart", C.c_void_p), ("gvec", C.c_void_p)]


# name -> (restype, argtypes); every symbol include/cpt_b200.h declares
_i, _ll, _sz, _p, _f = C.c_int, C.c_longlong, C.c_size_t, C.c_void_p, C.c_float
SYMBOLS = {
    "cpt_last_error": (C.c_char_p, []),
    "cpt_abi_version": (_i, []),
    "cpt_create": (_i, [C.POINTER(Config), _i, C.POINTER(_p)]),
    "cpt_destroy": (_i, [_p]),
    "cpt_set_weights": (_i, [_p, C.POINTER(Weights), _p]),
    "cpt_workspace_bytes": (_sz, [_p, _i, _i, _i]),
    "cpt_encoder_forward": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _sz, _p, _p, _p]),
    "cpt_mlm_gather_forward": (_i, [_p, _p, _p, _i, _i, _p, _p, _i, _p, _sz, _p]),
    "cpt_mlm_scores_forward": (_i, [_p, _p, _p, _ll, _p, _sz, _p]),
    "cpt_mlm_scores_workspace_bytes": (_sz, [_p, _ll]),
    "cpt_nsp_forward": (_i, [_p, _p, _p, _i, _p]),
    "cpt_head_linear": (_i, [_p, _p, _p, _i, _p, _p, _i, _p]),
    "cpt_exchange_create": (_i, [_i, _i, _i, _i, _i, C.POINTER(_p), _p]),
    "cpt_exchange_connect": (_i, [_p, _p]),
    "cpt_exchange_destroy": (_i, [_p]),
    "cpt_exchange_rows": (_i, [_p, _p, _p, _p, _i, _p]),
    "cpt_mlm_gather_exchange": (_i, [_p, _p, _p, _p, _i, _i, _p, _p, _i, _p, _sz, _p]),
    "cpt_train_enable": (_i, [_p, _i]),
    "cpt_train_set_progress_callback": (_i, [_p, _p, _p]),
    "cpt_train_tape_bytes": (_sz, [_p, _i, _i, _i, _i]),
    "cpt_train_forward_mlm": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _i, C.POINTER(Dropout), _p, _sz,
                                   _p]),
    "cpt_train_backward_mlm": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _i, C.POINTER(Dropout), _p, _p, _sz,
                                    C.POINTER(Grads)]),
    "cpt_train_forward_nsp": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _i, C.POINTER(Dropout), _p, _sz,
                                   _p]),
    "cpt_train_backward_nsp": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _i, C.POINTER(Dropout), _p, _p, _sz,
                                    C.POINTER(Grads)]),
    "cpt_adamw_step": (_i, [_i, _p, _p, _p, _i, _f, _f, _f, _i, _p]),
    "cpt_grad_clip_scale": (_i, [_i, _p, _p, _p, _i, _f, _p, _p, _p, _p]),
    "cpt_assemble_inputs": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "cpt_score_queries": (_i, [_p, _p, _p, _ll, _i, _i, _p, _p, _p, _p, _i, _p, _p, _p, _p]),
    "cpt_check_async_error": (_i, [_p, _p]),
    "cpt_kernel_name": (C.c_char_p, [_i]),
    "cpt_launch_count": (_ll, [_p]),
    "cpt_profile_enable": (_i, [_p, _i]),
    "cpt_profile_read": (_i, [_p, C.POINTER(C.c_double), C.POINTER(_ll)]),
    "cpt_gemm_trace": (_i, [_p, C.POINTER(_ll), _i]),
    "cpt_gemm": (_i, [_p, _p, _p, _ll, _p, _ll, _i, _i, _i, _p, _p, _ll, _i, _i, _p, _ll, _i]),
    "cpt_chain_run": (_i, [_p, _p, C.POINTER(ChainStage), _i]),
    "cpt_chain_trace": (_i, [_p, C.POINTER(_ll), _ll, C.POINTER(_i), C.POINTER(_i)]),
    "cpt_attention": (_i, [_p, _p, _p, _p, _i, _i, _p, _i]),
    "cpt_attention_backward": (_i, [_p, _p, _p, _p, _p, _i, _i, _p, _i]),
    "cpt_layernorm": (_i, [_p, _p, _p, _i, _p, _p, _f, _p, _p]),
    "cpt_cast16": (_i, [_p, _p, _p, _ll, _i, _i, _p]),
}

PROGRESS_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int)  # cpt_progress_fn

_lib = None


def load():
    """Load (once) and return the ctypes library with typed entry points."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CptError("cpt_b200: %s is missing — build it (`make -C cpt_b200/csrc`); there is no fallback path"
                       % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.cpt_abi_version() != ABI_VERSION:
        raise CptError("cpt_b200: ABI version mismatch (library %d, binding %d) — rebuild"
                       % (lib.cpt_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().cpt_last_error()
        raise CptError("cpt_b200: " + (msg.decode("utf-8", "replace") if msg else "error %d" % rc))
