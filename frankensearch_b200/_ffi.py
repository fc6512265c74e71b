"""ctypes binding of libfsgpu.so (include/fsgpu.h).

The library is the product; this module only declares its C ABI.  There is no fallback: if the
shared object is missing the import fails loudly, and every entry point fails with
`SearchError(kind="SubsystemError")` when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfsgpu.so")

# fsgpu_status -> SearchError variant (crates/frankensearch-core/src/error.rs:12-247)
STATUS_NAMES = {
    0: "Ok",
    1: "DimensionMismatch",
    2: "InvalidConfig",
    3: "IndexCorrupted",
    4: "EmbeddingFailed",
    5: "Cancelled",
    6: "SubsystemError",
    7: "Io",
}


ABI_VERSION = 4  # FSGPU_ABI_VERSION of include/fsgpu.h this binding was written against


class SearchError(Exception):
    """Mirror of `SearchError` (crates/frankensearch-core/src/error.rs:12)."""

    def __init__(self, kind: str, message: str):
        super().__init__(f"{kind}: {message}")
        self.kind = kind
        self.message = message


class Hit(C.Structure):
    _fields_ = [("row", C.c_uint32), ("score", C.c_float)]


class FusedHitC(C.Structure):
    _fields_ = [
        ("rrf_score", C.c_double),
        ("semantic_rank", C.c_int32),
        ("lexical_rank", C.c_int32),
        ("semantic_row", C.c_uint32),
        ("semantic_score", C.c_float),
        ("lexical_score", C.c_float),
        ("in_both_sources", C.c_uint32),
    ]


class IndexOptions(C.Structure):
    _fields_ = [
        ("device", C.c_int32),
        ("reduce_order", C.c_int32),
        ("tail_fma", C.c_int32),
        ("slab_is_device", C.c_int32),
        ("row_base", C.c_uint64),
        ("int8_codes", C.c_int32),
        ("flags", C.c_int32),
    ]


class Profile(C.Structure):
    _fields_ = [("scan_launches", C.c_uint64), ("merge_launches", C.c_uint64), ("other_launches", C.c_uint64),
                ("scan_bytes", C.c_uint64), ("scan_ms", C.c_double), ("mma_launches", C.c_uint64),
                ("mma_flops", C.c_double), ("redo_queries", C.c_uint64), ("i8_launches", C.c_uint64),
                ("pair_launches", C.c_uint64), ("quad_launches", C.c_uint64)]


class MiniLmLayerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "qkv_w", "qkv_b", "attn_out_w", "attn_out_b", "attn_ln_g", "attn_ln_b", "ffn_in_w", "ffn_in_b",
        "ffn_out_w", "ffn_out_b", "ffn_ln_g", "ffn_ln_b")]


class MiniLmWeights(C.Structure):
    _fields_ = [("vocab_size", C.c_uint32), ("max_positions", C.c_uint32), ("n_layers", C.c_uint32),
                ("hidden", C.c_uint32), ("heads", C.c_uint32), ("intermediate", C.c_uint32),
                ("ln_eps", C.c_float), ("reserved", C.c_uint32),
                ("word_emb", C.c_void_p), ("pos_emb", C.c_void_p), ("type_emb", C.c_void_p),
                ("emb_ln_g", C.c_void_p), ("emb_ln_b", C.c_void_p), ("layers", C.POINTER(MiniLmLayerWeights))]


class MiniLmProfile(C.Structure):
    _fields_ = [("gemm_launches", C.c_uint64), ("other_launches", C.c_uint64), ("gemm_flops", C.c_double),
                ("gemm_ms", C.c_double)]


class RrfConfigC(C.Structure):
    _fields_ = [
        ("k", C.c_double),
        ("lexical_weight", C.c_double),
        ("semantic_weight", C.c_double),
        ("tiebreak", C.c_int32),
        ("reserved", C.c_int32),
    ]


# Every symbol include/fsgpu.h declares (tests/test_abi.py checks the header against this list).
EXPORTS = [
    "fsgpu_abi_version", "fsgpu_last_error", "fsgpu_device_count", "fsgpu_index_options_default",
    "fsgpu_index_create_f16", "fsgpu_index_create_f32", "fsgpu_index_open_fsvi", "fsgpu_index_destroy",
    "fsgpu_index_rows", "fsgpu_index_dim", "fsgpu_index_row_base", "fsgpu_index_device",
    "fsgpu_index_device_slab", "fsgpu_index_set_doc_ids", "fsgpu_index_doc_id",
    "fsgpu_index_set_tombstones", "fsgpu_index_read_tombstones", "fsgpu_index_int8_ready",
    "fsgpu_index_read_codes_i8",
    "fsgpu_index_set_wal", "fsgpu_index_wal_rows", "fsgpu_index_zero_signal_state", "fsgpu_index_read_rows_f16", "fsgpu_index_profile_enable",
    "fsgpu_index_profile_read", "fsgpu_index_last_status", "fsgpu_measure_tensor_peak", "fsgpu_search_top_k", "fsgpu_search_top_k_device",
    "fsgpu_search_top_k_filtered", "fsgpu_search_top_k_filtered_device",
    "fsgpu_index_set_doc_hashes", "fsgpu_search_top_k_hashes",
    "fsgpu_search_top_k_two_pass", "fsgpu_index_read_two_pass_codes",
    "fsgpu_merge_top_k_device", "fsgpu_merge_top_k_hits_device", "fsgpu_scores_for_rows", "fsgpu_scores_for_rows_device",
    "fsgpu_scores_for_hits_device", "fsgpu_merge_payload_device",
    "fsgpu_rrf_fuse", "fsgpu_rrf_fuse_device", "fsgpu_blend_two_tier", "fsgpu_blend_two_tier_device", "fsgpu_potion_create",
    "fsgpu_potion_destroy", "fsgpu_potion_embed", "fsgpu_potion_embed_device",
    "fsgpu_synth_rows_device",
    "fsgpu_sharded_create_f16", "fsgpu_sharded_from_shards", "fsgpu_sharded_destroy", "fsgpu_sharded_shard_count",
    "fsgpu_sharded_rows", "fsgpu_sharded_shard", "fsgpu_sharded_is_direct", "fsgpu_sharded_search_top_k",
    "fsgpu_minilm_create", "fsgpu_minilm_load", "fsgpu_minilm_destroy", "fsgpu_minilm_embed", "fsgpu_minilm_embed_device",
    "fsgpu_minilm_profile_enable", "fsgpu_minilm_profile_read",
]

_vp = C.c_void_p
_lib = None


def lib() -> C.CDLL:
    """Load libfsgpu.so (built in-tree by `make -C frankensearch_b200/csrc` / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C frankensearch_b200/csrc` "
            "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.fsgpu_abi_version.restype = C.c_int
    L.fsgpu_last_error.restype = C.c_char_p
    L.fsgpu_device_count.argtypes = [C.POINTER(C.c_int)]
    L.fsgpu_index_options_default.argtypes = [C.POINTER(IndexOptions)]
    L.fsgpu_index_options_default.restype = None
    L.fsgpu_index_create_f16.argtypes = [_vp, C.c_uint64, C.c_uint32, _vp, C.POINTER(IndexOptions), C.POINTER(_vp)]
    L.fsgpu_index_create_f32.argtypes = [_vp, C.c_uint64, C.c_uint32, _vp, C.POINTER(IndexOptions), C.POINTER(_vp)]
    L.fsgpu_index_open_fsvi.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.POINTER(IndexOptions), C.POINTER(_vp)]
    L.fsgpu_index_destroy.argtypes = [_vp]
    L.fsgpu_index_destroy.restype = None
    L.fsgpu_index_rows.argtypes = [_vp]
    L.fsgpu_index_rows.restype = C.c_uint64
    L.fsgpu_index_dim.argtypes = [_vp]
    L.fsgpu_index_dim.restype = C.c_uint32
    L.fsgpu_index_row_base.argtypes = [_vp]
    L.fsgpu_index_row_base.restype = C.c_uint64
    L.fsgpu_index_device.argtypes = [_vp]
    L.fsgpu_index_device.restype = C.c_int
    L.fsgpu_index_device_slab.argtypes = [_vp]
    L.fsgpu_index_device_slab.restype = _vp
    L.fsgpu_index_set_doc_ids.argtypes = [_vp, _vp, _vp]
    L.fsgpu_index_doc_id.argtypes = [_vp, C.c_uint64, C.POINTER(_vp), C.POINTER(C.c_uint32)]
    L.fsgpu_index_set_tombstones.argtypes = [_vp, _vp]
    L.fsgpu_index_read_tombstones.argtypes = [_vp, _vp]
    L.fsgpu_index_int8_ready.argtypes = [_vp]
    L.fsgpu_index_read_codes_i8.argtypes = [_vp, C.c_uint64, C.c_uint64, _vp, C.POINTER(C.c_float)]
    L.fsgpu_index_set_doc_hashes.argtypes = [_vp, _vp]
    L.fsgpu_search_top_k_hashes.argtypes = [_vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp, C.c_uint32, _vp,
                                            _vp, _vp, C.POINTER(C.c_int)]
    L.fsgpu_index_zero_signal_state.argtypes = [_vp, _vp]
    L.fsgpu_search_top_k_two_pass.argtypes = [_vp, _vp, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, _vp, _vp]
    L.fsgpu_index_read_two_pass_codes.argtypes = [_vp, C.c_int, _vp]
    L.fsgpu_index_set_wal.argtypes = [_vp, _vp, C.c_uint32, C.c_uint64]
    L.fsgpu_index_wal_rows.argtypes = [_vp]
    L.fsgpu_index_wal_rows.restype = C.c_uint32
    L.fsgpu_index_read_rows_f16.argtypes = [_vp, C.c_uint64, C.c_uint64, _vp]
    L.fsgpu_index_profile_enable.argtypes = [_vp, C.c_int]
    L.fsgpu_index_profile_read.argtypes = [_vp, C.POINTER(Profile), C.c_int]
    L.fsgpu_index_last_status.argtypes = [_vp, _vp]
    L.fsgpu_measure_tensor_peak.argtypes = [C.c_int, C.c_int, C.c_uint32, C.POINTER(C.c_double)]
    L.fsgpu_search_top_k.argtypes = [_vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp, _vp]
    L.fsgpu_search_top_k_device.argtypes = [_vp, _vp, C.c_uint32, C.c_uint32, _vp, _vp, _vp, _vp]
    L.fsgpu_search_top_k_filtered.argtypes = [_vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp, _vp, _vp]
    L.fsgpu_search_top_k_filtered_device.argtypes = [_vp, _vp, C.c_uint32, C.c_uint32, _vp, _vp, _vp, _vp, _vp]
    L.fsgpu_merge_top_k_device.argtypes = [C.c_int, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64,
                                           C.c_uint64, C.c_uint32, _vp, _vp, _vp, _vp]
    L.fsgpu_merge_top_k_hits_device.argtypes = [C.c_int, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64,
                                                C.c_uint64, C.c_uint32, _vp, _vp, _vp, _vp]
    L.fsgpu_scores_for_rows.argtypes = [_vp, _vp, C.c_uint32, _vp, C.c_uint32, _vp, _vp]
    L.fsgpu_scores_for_rows_device.argtypes = [_vp, _vp, C.c_uint32, _vp, C.c_uint32, _vp, _vp, _vp]
    L.fsgpu_rrf_fuse.argtypes = [C.c_int, C.POINTER(RrfConfigC), C.c_uint32, _vp, _vp, _vp, _vp, C.c_uint32,
                                 _vp, _vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp, _vp]
    L.fsgpu_rrf_fuse_device.argtypes = [C.c_int, C.POINTER(RrfConfigC), C.c_uint32, _vp, _vp, _vp, _vp,
                                        C.c_uint32, _vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp,
                                        _vp, _vp]
    L.fsgpu_blend_two_tier.argtypes = [C.c_int, C.c_float, _vp, _vp, _vp, C.c_uint32, _vp, _vp, _vp, _vp,
                                       C.c_uint32, _vp, C.POINTER(C.c_uint32)]
    L.fsgpu_blend_two_tier_device.argtypes = [C.c_int, C.c_float, C.c_uint32, _vp, _vp, _vp, C.c_uint32, _vp, _vp, _vp,
                                              _vp, _vp, C.c_uint32, _vp, _vp, _vp]
    L.fsgpu_scores_for_hits_device.argtypes = [_vp, _vp, C.c_uint32, _vp, C.c_uint32, _vp, _vp, _vp]
    L.fsgpu_merge_payload_device.argtypes = [C.c_int, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64,
                                             C.c_uint64, C.c_uint64, C.c_uint64, _vp, C.c_uint32, _vp, _vp, _vp]
    L.fsgpu_potion_create.argtypes = [_vp, C.c_uint64, C.c_uint32, C.c_int, C.POINTER(_vp)]
    L.fsgpu_potion_destroy.argtypes = [_vp]
    L.fsgpu_potion_destroy.restype = None
    L.fsgpu_potion_embed.argtypes = [_vp, _vp, _vp, C.c_uint32, _vp]
    L.fsgpu_potion_embed_device.argtypes = [_vp, _vp, _vp, C.c_uint32, _vp, _vp]
    L.fsgpu_synth_rows_device.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32,
                                          C.c_uint32, C.c_float, _vp, _vp]
    L.fsgpu_sharded_create_f16.argtypes = [_vp, C.c_uint64, C.c_uint32, _vp, _vp, C.c_int, C.POINTER(IndexOptions), C.POINTER(_vp)]
    L.fsgpu_sharded_from_shards.argtypes = [_vp, C.c_int, C.c_int, C.POINTER(_vp)]
    L.fsgpu_sharded_destroy.argtypes = [_vp]
    L.fsgpu_sharded_destroy.restype = None
    L.fsgpu_sharded_shard_count.argtypes = [_vp]
    L.fsgpu_sharded_rows.argtypes = [_vp]
    L.fsgpu_sharded_rows.restype = C.c_uint64
    L.fsgpu_sharded_shard.argtypes = [_vp, C.c_int]
    L.fsgpu_sharded_shard.restype = _vp
    L.fsgpu_sharded_is_direct.argtypes = [_vp, C.c_int]
    L.fsgpu_sharded_search_top_k.argtypes = [_vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp, _vp]
    L.fsgpu_minilm_create.argtypes = [C.POINTER(MiniLmWeights), C.c_int, C.POINTER(_vp)]
    L.fsgpu_minilm_load.argtypes = [C.c_char_p, C.c_int, C.POINTER(_vp)]
    L.fsgpu_minilm_destroy.argtypes = [_vp]
    L.fsgpu_minilm_destroy.restype = None
    L.fsgpu_minilm_embed.argtypes = [_vp, _vp, _vp, C.c_uint32, C.c_uint32, _vp]
    L.fsgpu_minilm_embed_device.argtypes = [_vp, _vp, _vp, C.c_uint32, C.c_uint32, _vp, _vp]
    L.fsgpu_minilm_profile_enable.argtypes = [_vp, C.c_int]
    L.fsgpu_minilm_profile_read.argtypes = [_vp, C.POINTER(MiniLmProfile), C.c_int]
    _lib = L
    return L


def check(status: int) -> None:
    if status != 0:
        msg = lib().fsgpu_last_error().decode("utf-8", "replace")
        raise SearchError(STATUS_NAMES.get(status, f"Status{status}"), msg)


def ptr(a):
    """Host pointer of a C-contiguous numpy array (None -> NULL)."""
    if a is None:
        return None
    return a.ctypes.data
