"""ctypes binding of the C ABI in include/mrag.h (libmrag.so).

This is the only way the Python shim reaches the GPU; if the library is missing or will not
load, importing the compute entry points fails loudly -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

# --- constants mirrored from include/mrag.h ---------------------------------------------------
MRAG_OK, MRAG_ERR_ARG, MRAG_ERR_CUDA, MRAG_ERR_OOM, MRAG_ERR_STATE = 0, -1, -2, -3, -4
MRAG_F32, MRAG_BF16 = 0, 1
MRAG_FUSED_K, MRAG_MAX_K = 128, 2048
MRAG_PAYER_WORDS, MRAG_SMALL_WORDS, MRAG_TAG_WORDS = 16, 4, 8
MRAG_CODE_NONE = 0xFFFF
F_PAYER, F_STATE, F_PROGRAM, F_AUTHORITY, F_SOURCE_TYPE = 1, 2, 4, 8, 16
F_DOC_EQ, F_DOC_POOL, F_TAG_STRICT, F_TAG_RELAXED, F_DOC_POOL_HANDLE = 32, 64, 128, 256, 512
MRAG_POOL_LEVELS = 4
OPT_DEVICE_IO, OPT_FORCE_GEMV, OPT_FORCE_MMA, OPT_NO_SYNC, OPT_FORCE_MMA128, OPT_COALESCE = 1, 2, 4, 8, 16, 32
MRAG_PHRASE_WORDS, MRAG_JPD_CATS, MRAG_JTAG_WORDS, MRAG_HYB_MAX_PHRASES = 2, 11, 4, 16
CF_SHORT_TEXT, CF_CONTACT_VALUE, CF_PROMOTED, CF_DTAG_OVERFLOW = 1, 2, 4, 8

EXPORTS = [
    "mrag_create", "mrag_destroy", "mrag_append", "mrag_append_device", "mrag_set_doc_tags",
    "mrag_tombstone_doc", "mrag_size", "mrag_capacity", "mrag_dim", "mrag_index_dtype", "mrag_device",
    "mrag_search", "mrag_set_row_base", "mrag_merge_topk", "mrag_filter_mask", "mrag_last_kernel_ms",
    "mrag_profile_begin", "mrag_profile_read", "mrag_launch_count", "mrag_last_scan_kind",
    "mrag_last_error", "mrag_version",
    "mrag_set_chunk_features", "mrag_set_doc_jtags", "mrag_search_hybrid", "mrag_dtag_mask", "mrag_exchange_merge", "mrag_save", "mrag_load",
    "mrag_set_row_ids", "mrag_set_dtag_overflow", "mrag_rerank_candidates", "mrag_pool_build", "mrag_pool_select", "mrag_pool_add_docs", "mrag_pool_docs", "mrag_pool_destroy",
    "mrag_set_prepared_event", "mrag_live_rows",
]


class RowMeta(C.Structure):
    """mrag_rowmeta (12 bytes)."""
    _fields_ = [
        ("doc_idx", C.c_uint32), ("payer", C.c_uint16), ("state", C.c_uint8), ("program", C.c_uint8),
        ("authority", C.c_uint8), ("source_type", C.c_uint8), ("valid", C.c_uint8), ("reserved", C.c_uint8),
    ]


class FilterStruct(C.Structure):
    """mrag_filter."""
    _fields_ = [
        ("flags", C.c_uint32),
        ("payer_any", C.c_uint64 * MRAG_PAYER_WORDS),
        ("payer_alt_any", C.c_uint64 * MRAG_PAYER_WORDS),
        ("alt_state", C.c_uint16), ("state_eq", C.c_uint16), ("program_eq", C.c_uint16),
        ("authority_eq", C.c_uint16), ("source_type_eq", C.c_uint16), ("reserved0", C.c_uint16),
        ("doc_eq", C.c_uint32),
        ("doc_pool", C.c_void_p), ("n_doc_pool", C.c_int64),
        ("tag_state_any", C.c_uint64 * MRAG_SMALL_WORDS),
        ("tag_program_any", C.c_uint64 * MRAG_SMALL_WORDS),
        ("tag_payer_any", C.c_uint64 * MRAG_PAYER_WORDS),
        ("tag_any", C.c_uint64 * MRAG_TAG_WORDS),
        ("pool", C.c_void_p),
    ]


class ChunkFeat(C.Structure):
    """mrag_chunkfeat (40 bytes)."""
    _fields_ = [
        ("phrase_bits", C.c_uint64 * MRAG_PHRASE_WORDS), ("jpd_hits", C.c_uint8 * MRAG_JPD_CATS), ("flags", C.c_uint8),
        ("length_score", C.c_float), ("dtags", C.c_uint16 * 4),
    ]


class PoolQuery(C.Structure):
    """mrag_pool_query."""
    _fields_ = [
        ("d_all", C.c_uint64 * MRAG_TAG_WORDS), ("p_all", C.c_uint64 * MRAG_TAG_WORDS),
        ("j_all", C.c_uint64 * 4), ("ahca", C.c_uint64 * 4),
        ("has_j", C.c_int32), ("has_d", C.c_int32), ("has_p", C.c_int32), ("has_ahca", C.c_int32),
    ]


class Candidate(C.Structure):
    """mrag_candidate (52 bytes)."""
    _fields_ = [("feat", ChunkFeat), ("sim", C.c_float), ("doc_idx", C.c_uint32), ("authority", C.c_uint8), ("dtag_match", C.c_uint8),
                ("reserved", C.c_uint8 * 2)]


class HybridQuery(C.Structure):
    """mrag_hybrid_query."""
    _fields_ = [
        ("n_phrases", C.c_int32),
        ("phrase_weight", C.c_float * MRAG_HYB_MAX_PHRASES),
        ("phrase_bit", C.c_int16 * MRAG_HYB_MAX_PHRASES),
        ("phrase_jbit", C.c_int16 * MRAG_HYB_MAX_PHRASES),
        ("phrase_dcode", C.c_uint16 * MRAG_HYB_MAX_PHRASES),
        ("qcat", C.c_float * MRAG_JPD_CATS),
        ("auth_score", C.c_float * 32),
        ("w_sim", C.c_float), ("w_auth", C.c_float), ("w_len", C.c_float), ("w_jpd", C.c_float), ("w_cov", C.c_float),
        ("boost", C.c_float), ("floor", C.c_float), ("contact_query", C.c_uint32),
        ("source_type_any", C.c_uint64 * MRAG_SMALL_WORDS),
    ]


class MragError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"mrag error {code}: {message}")
        self.code = code


_lib = None


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True):
    """Load libmrag.so (building it first if the sources are newer and nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    override = os.environ.get("MRAG_LIB")          # A/B experiments: load exactly this build
    if override:
        _build.LIB = override
    elif build_if_missing and _build.nvcc_path() is not None:
        _build.build_lib()
    if not os.path.exists(_build.LIB):
        raise ImportError(
            f"{_build.LIB} is missing: build it with `python -m mobius-rag_b200.build` "
            "(__graft_entry__.build()). There is no CPU fallback for the scan.")
    lib = C.CDLL(_build.LIB)
    vp, i64, i32, u32, f32p = C.c_void_p, C.c_int64, C.c_int, C.c_uint32, C.POINTER(C.c_float)
    lib.mrag_create.restype = i32
    lib.mrag_create.argtypes = [C.POINTER(vp), i32, i32, i32, i64]
    lib.mrag_destroy.restype = i32
    lib.mrag_destroy.argtypes = [vp]
    lib.mrag_append.restype = i32
    lib.mrag_append.argtypes = [vp, vp, i64, vp, C.POINTER(i64)]
    lib.mrag_append_device.restype = i32
    lib.mrag_append_device.argtypes = [vp, vp, i64, vp, C.POINTER(i64), vp]
    lib.mrag_set_doc_tags.restype = i32
    lib.mrag_set_doc_tags.argtypes = [vp, i64, vp, i64]
    lib.mrag_tombstone_doc.restype = i32
    lib.mrag_tombstone_doc.argtypes = [vp, u32, C.POINTER(i64)]
    for name in ("mrag_size", "mrag_capacity"):
        getattr(lib, name).restype = i64
        getattr(lib, name).argtypes = [vp]
    for name in ("mrag_dim", "mrag_index_dtype", "mrag_device"):
        getattr(lib, name).restype = i32
        getattr(lib, name).argtypes = [vp]
    lib.mrag_search.restype = i32
    lib.mrag_search.argtypes = [vp, vp, i32, i32, C.POINTER(FilterStruct), vp, vp, vp, u32, vp]
    lib.mrag_set_row_base.restype = i32
    lib.mrag_set_row_base.argtypes = [vp, i64]
    lib.mrag_live_rows.restype = i32
    lib.mrag_live_rows.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.mrag_set_prepared_event.restype = i32
    lib.mrag_set_prepared_event.argtypes = [vp, vp]
    lib.mrag_set_row_ids.restype = i32
    lib.mrag_set_row_ids.argtypes = [vp, i64, vp, i64]
    lib.mrag_rerank_candidates.restype = i32
    lib.mrag_rerank_candidates.argtypes = [vp, vp, i64, C.POINTER(HybridQuery), vp, vp, vp]
    lib.mrag_set_dtag_overflow.restype = i32
    lib.mrag_set_dtag_overflow.argtypes = [vp, vp, vp, i64]
    lib.mrag_pool_build.restype = i32
    lib.mrag_pool_build.argtypes = [vp, C.POINTER(PoolQuery), C.POINTER(vp), C.POINTER(i64)]
    lib.mrag_pool_select.restype = i32
    lib.mrag_pool_select.argtypes = [vp, i32, i64, C.POINTER(i64)]
    lib.mrag_pool_add_docs.restype = i32
    lib.mrag_pool_add_docs.argtypes = [vp, vp, i64]
    lib.mrag_pool_docs.restype = i32
    lib.mrag_pool_docs.argtypes = [vp, vp, i64, C.POINTER(i64)]
    lib.mrag_pool_destroy.restype = i32
    lib.mrag_pool_destroy.argtypes = [vp]
    lib.mrag_merge_topk.restype = i32
    lib.mrag_merge_topk.argtypes = [i32, i32, i32, i32, vp, vp, vp, i64, i64, i64, vp, vp, vp, vp]
    lib.mrag_filter_mask.restype = i32
    lib.mrag_filter_mask.argtypes = [vp, C.POINTER(FilterStruct), vp, C.POINTER(i64), vp]
    lib.mrag_last_kernel_ms.restype = C.c_float
    lib.mrag_last_kernel_ms.argtypes = [i32]
    lib.mrag_profile_begin.restype = i32
    lib.mrag_profile_begin.argtypes = [i32]
    lib.mrag_profile_read.restype = i32
    lib.mrag_profile_read.argtypes = [i32, f32p, i32]
    lib.mrag_launch_count.restype = i64
    lib.mrag_launch_count.argtypes = []
    lib.mrag_set_chunk_features.restype = i32
    lib.mrag_set_chunk_features.argtypes = [vp, i64, vp, i64]
    lib.mrag_set_doc_jtags.restype = i32
    lib.mrag_set_doc_jtags.argtypes = [vp, i64, vp, i64]
    lib.mrag_search_hybrid.restype = i32
    lib.mrag_search_hybrid.argtypes = [vp, vp, i32, i32, C.POINTER(FilterStruct), vp, vp, vp, vp, vp, vp]
    lib.mrag_dtag_mask.restype = i32
    lib.mrag_dtag_mask.argtypes = [vp, C.POINTER(FilterStruct), vp, i32, vp, vp]
    lib.mrag_save.restype = i32
    lib.mrag_save.argtypes = [vp, C.c_char_p, i64]
    lib.mrag_load.restype = i32
    lib.mrag_load.argtypes = [C.POINTER(vp), C.c_char_p, i32, i64, C.POINTER(i64)]
    lib.mrag_exchange_merge.restype = i32
    lib.mrag_exchange_merge.argtypes = [i32, i32, i32, i32, i32, vp, i64, i64, i64, u32, vp, vp, vp, vp]
    for name in ("mrag_last_scan_kind", "mrag_last_error", "mrag_version"):
        getattr(lib, name).restype = C.c_char_p
        getattr(lib, name).argtypes = []
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != MRAG_OK:
        raise MragError(rc, load().mrag_last_error().decode("utf-8", "replace"))
