"""ctypes binding of libivlm_b200.so (the C ABI declared in include/ivlm_b200.h).

There is no fallback: if the shared library is missing or a symbol is absent, importing the product
path raises.  The oracle under oracle/ is test infrastructure and is never imported from here.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libivlm_b200.so"
HEADER = PKG.parent / "include" / "ivlm_b200.h"

BF16, F32, I32, I64 = 0, 1, 2, 3
ACT_NONE, ACT_GELU, ACT_QUICK_GELU, ACT_RELU, ACT_SILU = 0, 1, 2, 3, 4
LIFT_HUMAN, LIFT_OBJECT_MESH, LIFT_POINTS = 0, 1, 2


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("lda", C.c_int64),
        ("w", C.c_void_p), ("ldw", C.c_int64),
        ("out", C.c_void_p), ("ldo", C.c_int64),
        ("bias", C.c_void_p),
        ("residual", C.c_void_p), ("ldr", C.c_int64),
        ("row_map", C.c_void_p),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("act", C.c_int32), ("out_dtype", C.c_int32), ("k_splits", C.c_int32),
        ("force_swap", C.c_int32), ("no_round", C.c_int32), ("res_row_mod", C.c_int32),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p),
        ("q_bs", C.c_int64), ("q_ts", C.c_int64), ("q_hs", C.c_int64),
        ("k_bs", C.c_int64), ("k_ts", C.c_int64), ("k_hs", C.c_int64),
        ("v_bs", C.c_int64), ("v_ts", C.c_int64), ("v_hs", C.c_int64),
        ("o_bs", C.c_int64), ("o_ts", C.c_int64), ("o_hs", C.c_int64),
        ("B", C.c_int32), ("H", C.c_int32), ("Sq", C.c_int32), ("Sk", C.c_int32), ("D", C.c_int32),
        ("scale", C.c_float), ("causal", C.c_int32),
        ("rel_h", C.c_void_p), ("rel_w", C.c_void_p),
        ("kh", C.c_int32), ("kw", C.c_int32),
    ]


class DecodeLinearArgs(C.Structure):
    """ivlm_decode_linear_args (include/ivlm_b200.h)."""
    _fields_ = [
        ("a", C.c_void_p), ("lda", C.c_int64),
        ("w", C.c_void_p), ("ldw", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("norm_gamma", C.c_void_p), ("norm_eps", C.c_float),
        ("epilogue", C.c_int32), ("act", C.c_int32),
        ("bias", C.c_void_p), ("residual", C.c_void_p), ("ldr", C.c_int64),
        ("out", C.c_void_p), ("ldo", C.c_int64), ("out_dtype", C.c_int32),
        ("positions", C.c_void_p), ("slot_map", C.c_void_p), ("cos_t", C.c_void_p), ("sin_t", C.c_void_p),
        ("k_cache", C.c_void_p), ("v_cache", C.c_void_p),
        ("H", C.c_int32), ("hd", C.c_int32), ("page_size", C.c_int32),
        ("prefetch_w", C.c_void_p), ("prefetch_ldw", C.c_int64),
        ("prefetch_N", C.c_int32), ("prefetch_K", C.c_int32), ("prefetch_stages", C.c_int32),
    ]


EPI_PLAIN, EPI_SWIGLU, EPI_ROPE_KV = 0, 1, 2


class WeightDesc(C.Structure):
    """ivlm_weight_desc."""
    _fields_ = [("name", C.c_char_p), ("ptr", C.c_void_p), ("dtype", C.c_int32), ("ndim", C.c_int32), ("shape", C.c_int64 * 4)]


class ModelDims(C.Structure):
    """ivlm_model_dims."""
    _fields_ = [("sam_img", C.c_int32), ("sam_patch", C.c_int32), ("sam_embed_dim", C.c_int32), ("sam_depth", C.c_int32),
                ("sam_heads", C.c_int32), ("sam_window", C.c_int32), ("sam_out_chans", C.c_int32), ("sam_global_mask", C.c_uint32),
                ("llm_hidden", C.c_int32), ("llm_intermediate", C.c_int32), ("llm_layers", C.c_int32), ("llm_heads", C.c_int32),
                ("llm_head_dim", C.c_int32), ("llm_vocab", C.c_int32), ("llm_rms_eps", C.c_float), ("llm_paired_layout", C.c_int32),
                ("clip_img", C.c_int32), ("clip_patch", C.c_int32), ("clip_hidden", C.c_int32), ("clip_heads", C.c_int32),
                ("clip_layers", C.c_int32), ("clip_ldk", C.c_int32), ("clip_eps", C.c_float)]


class ClipEncodeArgs(C.Structure):
    """ivlm_clip_encode_args."""
    _fields_ = [("images", C.c_void_p), ("feats", C.c_void_p), ("B", C.c_int32), ("patch_rows", C.c_void_p), ("cls_rows", C.c_void_p),
                ("arena", C.c_void_p), ("arena_bytes", C.c_size_t)]


class MaskDecodeArgs(C.Structure):
    """ivlm_mask_decode_args."""
    _fields_ = [("emb", C.c_void_p), ("prompt", C.c_void_p), ("lowres", C.c_void_p), ("n", C.c_int32), ("V", C.c_int32), ("heads", C.c_int32),
                ("tok_idx", C.c_void_p), ("arena", C.c_void_p), ("arena_bytes", C.c_size_t)]


class SamEncodeArgs(C.Structure):
    """ivlm_sam_encode_args."""
    _fields_ = [("images", C.c_void_p), ("emb", C.c_void_p), ("N", C.c_int32), ("win_map", C.c_void_p), ("win_inv", C.c_void_p),
                ("win_pads", C.c_void_p), ("n_pads", C.c_int32), ("arena", C.c_void_p), ("arena_bytes", C.c_size_t)]


class LlmPrefillArgs(C.Structure):
    """ivlm_llm_prefill_args."""
    _fields_ = [("embeds", C.c_void_p), ("positions", C.c_void_p), ("slot_map", C.c_void_p), ("k_cache", C.POINTER(C.c_void_p)),
                ("v_cache", C.POINTER(C.c_void_p)), ("hidden", C.c_void_p), ("next_tok", C.c_void_p), ("last_rows", C.c_void_p),
                ("B", C.c_int32), ("S", C.c_int32), ("max_len", C.c_int32), ("page_size", C.c_int32), ("arena", C.c_void_p),
                ("arena_bytes", C.c_size_t)]


class LlmDecodeArgs(C.Structure):
    """ivlm_llm_decode_args."""
    _fields_ = [("state", C.c_void_p), ("S", C.c_int32), ("S_rows", C.c_void_p), ("scripted", C.c_void_p), ("G", C.c_int32),
                ("next", C.c_void_p), ("done", C.c_void_p), ("out_tokens", C.c_void_p), ("tok", C.c_void_p), ("pos", C.c_void_p),
                ("slot", C.c_void_p), ("seq_lens", C.c_void_p), ("slot_base", C.c_void_p), ("eos", C.c_int32), ("pad", C.c_int32),
                ("B", C.c_int32), ("k_cache", C.POINTER(C.c_void_p)), ("v_cache", C.POINTER(C.c_void_p)), ("block_table", C.c_void_p),
                ("max_pages", C.c_int32), ("page_size", C.c_int32), ("hidden", C.c_void_p), ("max_len", C.c_int32),
                ("hid_step", C.c_void_p), ("arena", C.c_void_p), ("arena_bytes", C.c_size_t)]


class RasterCam(C.Structure):
    """ivlm_raster_cam (include/ivlm_b200.h)."""
    _fields_ = [("R", C.c_float * 9), ("T", C.c_float * 3), ("C", C.c_float * 3), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("z_clip", C.c_float)]


RASTER_MAX_VIEWS = 8


def declared_symbols() -> list[str]:
    """Every function the public header declares (used by the CPU-side ABI test)."""
    txt = HEADER.read_text()
    return sorted(set(re.findall(r"IVLM_API\s+[\w\s\*]+?\b(ivlm_\w+)\s*\(", txt)))


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m interactvlm_b200.build` "
                "(there is no CPU or PyTorch fallback for the hot path)")
        _lib = C.CDLL(str(LIB_PATH))
        _lib.ivlm_last_error.restype = C.c_char_p
        _lib.ivlm_launch_count.restype = C.c_uint64
        _lib.ivlm_lift_nnz.restype = C.c_int64
        _lib.ivlm_sam_encode_arena_bytes.restype = C.c_size_t
        _lib.ivlm_llm_arena_bytes.restype = C.c_size_t
        _lib.ivlm_clip_encode_arena_bytes.restype = C.c_size_t
        _lib.ivlm_mask_decode_arena_bytes.restype = C.c_size_t
        for name in declared_symbols():
            if not hasattr(_lib, name):
                raise RuntimeError(f"libivlm_b200.so does not export {name}")
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().ivlm_last_error().decode(errors="replace")
        raise RuntimeError(f"ivlm_b200 {what} failed ({status}): {msg}")
