"""ctypes binding of ``libmodcr_b200.so`` (C ABI declared in ``include/modcr_b200.h``).

The library is the product: there is no Python / PyTorch fallback.  If the shared
object is missing, or the device is not an sm_100 GPU, every compute call raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MODCR_B200_LIB: developer override (e.g. the `make trace` build with kernel timelines compiled in)
LIB_PATH = os.environ.get("MODCR_B200_LIB") or os.path.join(_HERE, "libmodcr_b200.so")

c_void_p, c_int, c_i32, c_i64, c_float, c_size_t = C.c_void_p, C.c_int, C.c_int32, C.c_int64, C.c_float, C.c_size_t


class GemmDesc(C.Structure):
    _fields_ = [("A", c_void_p), ("B", c_void_p), ("C", c_void_p),
                ("lda", c_i64), ("ldb", c_i64), ("ldc", c_i64),
                ("a_mn", c_i32), ("b_mn", c_i32), ("c_fp32", c_i32), ("accumulate", c_i32), ("split_k", c_i32),
                ("M", c_i32), ("N", c_i32), ("K", c_i32),
                ("alpha", c_float), ("bias", c_void_p), ("act", c_i32),
                ("aux", c_void_p), ("ldaux", c_i64), ("epi_in", c_void_p), ("ld_in", c_i64), ("in_op", c_i32),
                ("a_colsum", c_void_p), ("drop_seed", c_void_p), ("drop_site", C.c_uint32), ("drop_p", c_float),
                ("in_mean", c_void_p), ("in_rstd", c_void_p), ("in_gamma", c_void_p), ("in_beta", c_void_p)]


class RowMap(C.Structure):
    _fields_ = [("rows_per_seq", c_i32), ("seq_stride", c_i32), ("row0", c_i32)]


class AttnDesc(C.Structure):
    _fields_ = [("N", c_i32), ("S", c_i32), ("P", c_i32), ("heads", c_i32), ("T", c_i32), ("mode", c_i32),
                ("scale", c_float),
                ("qkv", c_void_p), ("pkv", c_void_p), ("kbias", c_void_p), ("cbits", c_void_p), ("cwords", c_i32),
                ("ctx", c_void_p), ("lse", c_void_p), ("probs_img", c_void_p),
                ("dctx", c_void_p), ("dalign", c_void_p), ("dqkv", c_void_p), ("dpkv", c_void_p), ("dq_ws", c_void_p),
                ("drop_seed", c_void_p), ("drop_site", C.c_uint32), ("drop_p", c_float), ("row_sel", c_void_p)]


class ScorerDesc(C.Structure):
    _fields_ = [("x", c_void_p), ("W", c_void_p), ("b", c_void_p),
                ("rows", c_i32), ("K", c_i32), ("C", c_i32), ("G", c_i32),
                ("hard", c_void_p), ("soft", c_void_p), ("logits", c_void_p), ("loss", c_void_p),
                ("mp", c_void_p), ("pred", c_void_p), ("matched", c_void_p), ("num_choice", c_i32),
                ("gout", c_float), ("dx", c_void_p), ("dW", c_void_p), ("db", c_void_p),
                ("loss_kind", c_i32)]


class PackDesc(C.Structure):
    _fields_ = [("header", c_void_p), ("tokens", c_void_p), ("feats", c_void_p),
                ("N", c_i32), ("T", c_i32), ("R", c_i32), ("D", c_i32),
                ("input_ids", c_void_p), ("token_type_ids", c_void_p), ("total_label", c_void_p),
                ("align_pos", c_void_p), ("label", c_void_p), ("label_f", c_void_p), ("input_mask", c_void_p),
                ("chunk_id", c_void_p), ("chunk_bits", c_void_p), ("cwords", c_i32), ("chunk_mask", c_void_p),
                ("img_feat", c_void_p)]


# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/modcr_b200.h
SIGNATURES = {
    "modcr_version": [],
    "modcr_last_error_string": [],
    "modcr_launch_count": [],
    "modcr_workspace_bytes": [c_int, C.POINTER(c_i64), c_int],
    "modcr_set_sm_limit": [c_i32],
    "modcr_gemm": [C.POINTER(GemmDesc), c_void_p],
    "modcr_layernorm_fwd": [c_void_p, c_i32, RowMap, c_void_p, c_void_p, c_void_p, c_void_p, RowMap, c_void_p, c_void_p,
                            c_i32, c_i32, c_float, c_void_p],
    "modcr_layernorm_bwd": [c_void_p, RowMap, c_void_p, c_i32, RowMap, c_void_p, c_void_p, c_void_p, c_void_p, RowMap,
                            c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_void_p, c_void_p, C.c_uint32, c_float, c_void_p],
    "modcr_txt_embed_ln_fwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, RowMap, c_void_p, c_void_p, c_i32, c_i32, c_i32, c_float, c_void_p],
    "modcr_txt_embed_ln_bwd": [c_void_p, RowMap, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i32,
                               c_i64, c_void_p],
    "modcr_img_embed_ln_fwd": [c_void_p, c_i64, c_i32, c_i32, c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_float,
                               c_void_p, c_void_p, c_void_p, c_void_p, RowMap, c_void_p, c_void_p, c_i32, c_void_p,
                               c_void_p],
    "modcr_img_embed_ln_bwd": [c_void_p, RowMap, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i32, c_i32, c_void_p],
    "modcr_prefix_attn_fwd": [C.POINTER(AttnDesc), c_void_p],
    "modcr_prefix_attn_bwd": [C.POINTER(AttnDesc), c_void_p],
    "modcr_chunk_mean_q": [c_void_p, c_i64, c_i32, c_i32, c_i32, c_void_p, c_i32, c_i32, c_void_p],
    "modcr_align_softmax_ce_fwd": [c_void_p, c_i32, c_i32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_i32, c_i32, c_i32, c_void_p],
    "modcr_align_softmax_ce_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_i32, c_i32, c_i32,
                                   c_void_p],
    "modcr_grad_sqnorm": [c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_void_p, c_void_p],
    "modcr_grad_clip": [c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_void_p, c_float, c_void_p],
    "modcr_adamw_step": [c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_void_p, c_void_p, c_float, c_float, c_float,
                         c_void_p],
    "modcr_multi_copy": [c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i32, c_float, c_void_p],
    "modcr_dropout": [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, C.c_uint32, c_float, c_void_p],
    "modcr_cross_attn_q1_fwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i32, c_i32, c_void_p,
                                C.c_uint32, c_float, c_void_p],
    "modcr_cross_attn_q1_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_i32, c_i32, c_i32, c_i32, c_void_p, C.c_uint32, c_float, c_void_p],
    "modcr_cls_pool_attn_fwd": [c_void_p, c_void_p, c_i32, c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i32, c_void_p,
                                C.c_uint32, c_float, c_void_p],
    "modcr_cls_pool_attn_bwd": [c_void_p, c_void_p, c_void_p, c_i32, c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i32,
                                c_void_p, C.c_uint32, c_float, c_void_p],
    "modcr_mc_scorer_ce_fwd": [C.POINTER(ScorerDesc), c_void_p],
    "modcr_mc_scorer_ce_bwd": [C.POINTER(ScorerDesc), c_void_p],
    "modcr_pack_batch": [C.POINTER(PackDesc), c_void_p],
    "modcr_cast_pad": [c_void_p, c_i64, c_void_p, c_i64, c_i64, c_i32, c_i64, c_void_p],
    "modcr_bf16_to_f32": [c_void_p, c_void_p, c_i64, c_i32, c_void_p],
    "modcr_colsum": [c_void_p, c_i64, c_i32, c_i32, c_void_p, c_void_p],
    "modcr_copy_rows": [c_void_p, c_i64, c_i32, c_void_p, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_void_p],
    "modcr_add_bf16": [c_void_p, c_void_p, c_void_p, c_i64, c_void_p],
    "modcr_relu_bwd": [c_void_p, c_void_p, c_void_p, c_i64, c_void_p],
    "modcr_rowln_fwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_float, c_void_p],
    "modcr_rowln_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i32,
                        c_void_p],
    "modcr_tanh_bwd": [c_void_p, c_void_p, c_void_p, c_i64, c_void_p],
    "modcr_pos_type_embed_fwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_i32, c_i32, c_float, c_void_p],
    "modcr_pos_type_embed_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_void_p],
    "modcr_rows_place": [c_void_p, c_void_p, c_void_p, RowMap, c_i32, c_i32, c_void_p],
    "modcr_rows_gather": [c_void_p, RowMap, c_void_p, c_i32, c_i32, c_void_p],
    "modcr_key_bias": [c_void_p, c_void_p, c_i64, c_void_p],
    "modcr_chunk_bits": [c_void_p, c_void_p, c_i32, c_i32, c_i32, c_void_p],
    "modcr_accum_f32_2d": [c_void_p, c_i64, c_void_p, c_i64, c_i64, c_i32, c_i32, c_void_p],
    "modcr_adamw_clip": [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_float, c_float, c_float, c_float, c_float,
                         c_float, c_float, c_void_p, c_float, c_void_p],
    "modcr_sumsq_f32": [c_void_p, c_i64, c_void_p, c_void_p],
}
_RESTYPES = {"modcr_last_error_string": C.c_char_p, "modcr_workspace_bytes": c_size_t,
             "modcr_launch_count": C.c_uint64}
OPTIONAL = {"modcr_adamw_clip", "modcr_sumsq_f32"}

_lib = None


def load():
    """Load the shared library (once) and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "modcr_b200: %s is missing -- build it with `make` (or `python -c 'import __graft_entry__ as g; "
            "g.build()'`); there is no CPU / PyTorch fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            if name in OPTIONAL:
                continue
            raise RuntimeError("modcr_b200: symbol %s missing from %s" % (name, LIB_PATH))
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_int)
    _lib = lib
    return lib


def set_sm_limit(sms):
    """SM budget of the persistent kernels (0 = all); returns the previous value."""
    return int(load().modcr_set_sm_limit(int(sms)))


def launch_count():
    """Kernels launched by libmodcr_b200.so in this process so far."""
    return int(load().modcr_launch_count())


def last_error():
    s = load().modcr_last_error_string()
    return s.decode() if s else ""


def check(rc, what):
    if rc != 0:
        raise RuntimeError("modcr_b200.%s failed (%d): %s" % (what, rc, last_error()))
