"""ctypes front-end of the CPU oracle (oracle/fs_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfs_oracle.so")

REDUCE_HALVES_PAIRWISE = 0
REDUCE_AVX_TREE = 1
REDUCE_HALVES_SEQUENTIAL = 2
REDUCE_HALVES_STRIDE2 = 3
REDUCE_SEQUENTIAL = 4
DEFAULT_REDUCE = REDUCE_HALVES_PAIRWISE

TIEBREAK_LEXICAL_THEN_ID = 0
TIEBREAK_HASH = 1


def build(force: bool = False) -> str:
    """Compile libfs_oracle.so with the committed Makefile (gcc only, a few seconds)."""
    src = os.path.join(_HERE, "fs_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libfs_oracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _declare(_lib)
    return _lib


_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)


def _declare(L: C.CDLL) -> None:
    L.fso_f16_to_f32.argtypes = [C.c_uint16]
    L.fso_f16_to_f32.restype = C.c_float
    L.fso_f32_to_f16.argtypes = [C.c_float]
    L.fso_f32_to_f16.restype = C.c_uint16
    L.fso_encode_f32_to_f16.argtypes = [_f32p, C.c_uint64, _u16p, C.c_int]
    L.fso_decode_f16_to_f32.argtypes = [_u16p, C.c_uint64, _f32p, C.c_int]
    L.fso_have_avx2.restype = C.c_int
    L.fso_dot_f16_f32.argtypes = [_u16p, _f32p, C.c_uint32, C.c_int, C.c_int, C.c_int]
    L.fso_dot_f16_f32.restype = C.c_float
    L.fso_search_top_k.argtypes = [_u16p, C.c_uint64, C.c_uint32, _u8p, _f32p, C.c_uint64, C.c_int,
                                   C.c_int, C.c_int, _u64p, _f32p]
    L.fso_search_top_k.restype = C.c_uint64
    L.fso_quantize_slab_i8.argtypes = [_u16p, C.c_uint64, C.c_void_p]
    L.fso_pack_slab_4bit.argtypes = [_u16p, C.c_uint64, C.c_uint32, _u8p]
    L.fso_search_two_pass.argtypes = [_u16p, C.c_uint64, C.c_uint32, _u8p, _f32p, C.c_uint64, C.c_uint64, C.c_int,
                                      C.c_int, C.c_int, _u64p, _f32p]
    L.fso_search_two_pass.restype = C.c_uint64
    L.fso_scores_for_rows.argtypes = [_u16p, C.c_uint64, C.c_uint32, _f32p, _u64p, C.c_uint64,
                                      C.c_int, C.c_int, _f32p, _u8p]
    L.fso_fnv1a64.argtypes = [_u8p, C.c_uint64]
    L.fso_fnv1a64.restype = C.c_uint64
    L.fso_rrf_fuse.argtypes = [_u8p, _u64p, _f32p, C.c_uint64, _u8p, _u64p, C.c_uint64, C.c_double,
                               C.c_double, C.c_double, C.c_int, C.c_uint64, C.c_uint64, C.c_int,
                               _i64p, _i64p, _f64p, _u8p]
    L.fso_rrf_fuse.restype = C.c_uint64
    L.fso_blend_two_tier.argtypes = [_u8p, _u64p, _u32p, _f32p, C.c_uint64, _u8p, _u64p, _u32p, _f32p,
                                     C.c_uint64, C.c_float, _u32p, _f32p, _i64p]
    L.fso_blend_two_tier.restype = C.c_uint64
    L.fso_blend_two_tier_aligned.argtypes = [_u8p, _u64p, _u32p, _f32p, C.c_uint64, _f32p, _u8p,
                                             C.c_float, _u32p, _f32p, _i64p]
    L.fso_blend_two_tier_aligned.restype = C.c_uint64
    L.fso_potion_embed.argtypes = [_f32p, C.c_uint64, C.c_uint32, _u32p, C.c_uint64, _f32p]
    L.fso_potion_embed.restype = C.c_uint32
    L.fso_l2_normalize.argtypes = [_f32p, C.c_uint32]
    L.fso_raw_vector.argtypes = [C.c_uint64, C.c_uint32, _f32p]
    L.fso_normalize.argtypes = [_f32p, C.c_uint32]
    L.fso_synth_rows.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32,
                                 C.c_float, C.c_int, _f32p, _u16p]


def _p(a: Optional[np.ndarray], ty):
    if a is None:
        return None
    return a.ctypes.data_as(ty)


def _c(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


# ─── conversions ────────────────────────────────────────────────────────────────────────────
def encode_f16(x, hw: bool = True) -> np.ndarray:
    """f32 -> f16 bit patterns (uint16), round-to-nearest-even (simd.rs:2245-2304)."""
    x = _c(x, np.float32)
    out = np.empty(x.shape, dtype=np.uint16)
    lib().fso_encode_f32_to_f16(_p(x, _f32p), x.size, _p(out, _u16p), int(hw))
    return out


def decode_f16(h, hw: bool = True) -> np.ndarray:
    h = _c(h, np.uint16)
    out = np.empty(h.shape, dtype=np.float32)
    lib().fso_decode_f16_to_f32(_p(h, _u16p), h.size, _p(out, _f32p), int(hw))
    return out


def dot_f16_f32(row_bits, query, reduce_order: int = DEFAULT_REDUCE, tail_fma: bool = True,
                impl: int = 1) -> np.float32:
    row_bits = _c(row_bits, np.uint16)
    query = _c(query, np.float32)
    assert row_bits.size == query.size
    return np.float32(lib().fso_dot_f16_f32(_p(row_bits, _u16p), _p(query, _f32p), query.size,
                                           reduce_order, int(tail_fma), impl))


# ─── scan + top-k ───────────────────────────────────────────────────────────────────────────
def search_top_k(slab_bits, query, limit: int, tombstones: Optional[np.ndarray] = None,
                 threads: Optional[int] = None, reduce_order: int = DEFAULT_REDUCE,
                 tail_fma: bool = True):
    """Exact top-`limit` (search.rs:426-494).  slab_bits: [n, dim] uint16.  Returns (rows u64, scores f32)."""
    slab_bits = _c(slab_bits, np.uint16)
    n, dim = slab_bits.shape
    query = _c(query, np.float32)
    assert query.size == dim, "DimensionMismatch"
    cap = min(int(limit), n)
    rows = np.empty(max(cap, 1), dtype=np.uint64)
    scores = np.empty(max(cap, 1), dtype=np.float32)
    bm = None if tombstones is None else _c(tombstones, np.uint8)
    got = lib().fso_search_top_k(_p(slab_bits, _u16p), n, dim, _p(bm, _u8p), _p(query, _f32p),
                                 int(limit), threads or host_threads(), reduce_order, int(tail_fma),
                                 _p(rows, _u64p), _p(scores, _f32p))
    return rows[:got].copy(), scores[:got].copy()


def quantize_slab_i8(slab_bits) -> np.ndarray:
    """quantize_f16_slab_to_i8_generic (simd.rs:1842-1859): [rows, dim] f16 bits -> int8 codes of the same shape."""
    s = _c(slab_bits, np.uint16)
    out = np.zeros(s.shape, dtype=np.int8)
    lib().fso_quantize_slab_i8(_p(s, _u16p), s.size, out.ctypes.data_as(C.c_void_p))
    return out


def pack_slab_4bit(slab_bits) -> np.ndarray:
    """pack_f16_slab_to_4bit_generic (simd.rs:2201-2233): [rows, dim] f16 bits -> [rows, ceil(dim / 2)] nibble bytes."""
    s = _c(slab_bits, np.uint16)
    n, dim = s.shape
    out = np.zeros((n, (dim + 1) // 2), dtype=np.uint8)
    lib().fso_pack_slab_4bit(_p(s, _u16p), n, dim, _p(out, _u8p))
    return out


def search_two_pass(slab_bits, query, k: int, candidate_multiplier: int, bits: int, tombstones: Optional[np.ndarray] = None,
                    reduce_order: int = DEFAULT_REDUCE, tail_fma: bool = True):
    """search_top_k_int8_two_pass (bits = 8, search.rs:571-650) / search_top_k_4bit_two_pass (bits = 4, :876-946) on a
    slab without WAL rows.  Returns (rows u64, scores f32)."""
    s = _c(slab_bits, np.uint16)
    n, dim = s.shape
    q = _c(query, np.float32)
    rows = np.zeros(max(min(k, n), 1), dtype=np.uint64)
    scores = np.zeros(max(min(k, n), 1), dtype=np.float32)
    tb = pack_bitmap(tombstones) if tombstones is not None else None
    got = lib().fso_search_two_pass(_p(s, _u16p), n, dim, _p(tb, _u8p), _p(q, _f32p), k, candidate_multiplier, bits,
                                    reduce_order, 1 if tail_fma else 0, _p(rows, _u64p), _p(scores, _f32p))
    return rows[:got], scores[:got]


def dot_f32_f32(a, b, reduce_order: int = DEFAULT_REDUCE) -> np.float32:
    """dot_product_f32_f32 (simd.rs:134-222, :1559-1587) — the score of a resident WAL row."""
    a, b = _c(a, np.float32), _c(b, np.float32)
    assert a.size == b.size, "DimensionMismatch"
    L = lib()
    L.fso_dot_f32_f32.restype = C.c_float
    L.fso_dot_f32_f32.argtypes = [_f32p, _f32p, C.c_uint32, C.c_int]
    return np.float32(L.fso_dot_f32_f32(_p(a, _f32p), _p(b, _f32p), a.size, reduce_order))


def search_top_k_wal(slab_bits, wal, query, limit: int, exclude: Optional[np.ndarray] = None,
                     wal_allow: Optional[np.ndarray] = None, threads: Optional[int] = None,
                     reduce_order: int = DEFAULT_REDUCE, tail_fma: bool = True):
    """search_top_k_internal with resident WAL rows (search.rs:426-494, :1449-1475).  slab_bits
    [n, dim] uint16, wal [n_wal, dim] float32; `exclude` / `wal_allow` are packed bitmaps.  Returns
    (rows u64 — WAL row w as n + w, scores f32), best-first, before the doc-id resolve step."""
    slab_bits = _c(slab_bits, np.uint16)
    n, dim = slab_bits.shape
    wal = _c(wal, np.float32).reshape(-1, dim)
    n_wal = wal.shape[0]
    query = _c(query, np.float32)
    assert query.size == dim, "DimensionMismatch"
    cap = max(min(int(limit), n + n_wal), 1)
    rows = np.empty(cap, dtype=np.uint64)
    scores = np.empty(cap, dtype=np.float32)
    ex = None if exclude is None else _c(exclude, np.uint8)
    wa = None if wal_allow is None else _c(wal_allow, np.uint8)
    L = lib()
    L.fso_search_top_k_wal.restype = C.c_uint64
    L.fso_search_top_k_wal.argtypes = [_u16p, C.c_uint64, C.c_uint32, _u8p, _f32p, C.c_uint64, _u8p, _f32p,
                                       C.c_uint64, C.c_int, C.c_int, C.c_int, _u64p, _f32p]
    got = L.fso_search_top_k_wal(_p(slab_bits, _u16p), n, dim, _p(ex, _u8p), _p(wal, _f32p), n_wal,
                                 _p(wa, _u8p), _p(query, _f32p), int(limit), threads or host_threads(),
                                 reduce_order, int(tail_fma), _p(rows, _u64p), _p(scores, _f32p))
    return rows[:got].copy(), scores[:got].copy()


def scores_for_rows(slab_bits, query, rows, reduce_order: int = DEFAULT_REDUCE, tail_fma: bool = True):
    slab_bits = _c(slab_bits, np.uint16)
    n, dim = slab_bits.shape
    query = _c(query, np.float32)
    rows = _c(rows, np.uint64)
    out = np.empty(rows.size, dtype=np.float32)
    present = np.empty(rows.size, dtype=np.uint8)
    lib().fso_scores_for_rows(_p(slab_bits, _u16p), n, dim, _p(query, _f32p), _p(rows, _u64p),
                              rows.size, reduce_order, int(tail_fma), _p(out, _f32p), _p(present, _u8p))
    return out, present.astype(bool)


def pack_bitmap(flags) -> np.ndarray:
    """bool[n] -> packed little-endian bitmap (bit r%8 of byte r//8)."""
    return np.packbits(np.asarray(flags, dtype=bool), bitorder="little")


def fnv1a64(data: bytes) -> int:
    buf = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(0, dtype=np.uint8)
    return int(lib().fso_fnv1a64(_p(_c(buf, np.uint8), _u8p), len(data)))


# ─── strings ────────────────────────────────────────────────────────────────────────────────
def pack_ids(ids: Sequence[str]):
    blobs = [s.encode("utf-8") if isinstance(s, str) else bytes(s) for s in ids]
    off = np.zeros(len(blobs) + 1, dtype=np.uint64)
    if blobs:
        off[1:] = np.cumsum([len(b) for b in blobs], dtype=np.uint64)
    data = np.frombuffer(b"".join(blobs) or b"\0", dtype=np.uint8).copy()
    return data, off


# ─── RRF ────────────────────────────────────────────────────────────────────────────────────
@dataclass
class FusedHit:
    doc_id: str
    rrf_score: float
    lexical_rank: Optional[int]
    semantic_rank: Optional[int]
    semantic_index: Optional[int]
    lexical_score: Optional[float]
    semantic_score: Optional[float]
    in_both_sources: bool


def rrf_fuse(lexical: Sequence[tuple], semantic: Sequence[tuple], limit: int, offset: int = 0,
             k: float = 60.0, lexical_weight: float = 1.0, semantic_weight: float = 1.0,
             tiebreak: int = TIEBREAK_LEXICAL_THEN_ID, dedup_semantic: bool = True):
    """lexical: [(doc_id, score)] in rank order; semantic: [(doc_id, index, score)] in rank order."""
    lb, lo = pack_ids([d for d, _ in lexical])
    ls = _c([s for _, s in lexical], np.float32)
    sb, so = pack_ids([d for d, _, _ in semantic])
    cap = max(1, int(limit))
    sem_pos = np.empty(cap, dtype=np.int64)
    lex_pos = np.empty(cap, dtype=np.int64)
    rrf = np.empty(cap, dtype=np.float64)
    both = np.empty(cap, dtype=np.uint8)
    n = lib().fso_rrf_fuse(_p(lb, _u8p), _p(lo, _u64p), _p(ls, _f32p), len(lexical), _p(sb, _u8p),
                           _p(so, _u64p), len(semantic), k, lexical_weight, semantic_weight, tiebreak,
                           int(limit), int(offset), int(dedup_semantic), _p(sem_pos, _i64p),
                           _p(lex_pos, _i64p), _p(rrf, _f64p), _p(both, _u8p))
    out = []
    for i in range(n):
        sp, lp = int(sem_pos[i]), int(lex_pos[i])
        doc = semantic[sp][0] if sp >= 0 else lexical[lp][0]
        out.append(FusedHit(doc, float(rrf[i]), lp if lp >= 0 else None, sp if sp >= 0 else None,
                            int(semantic[sp][1]) if sp >= 0 else None,
                            float(np.float32(lexical[lp][1])) if lp >= 0 else None,
                            float(np.float32(semantic[sp][2])) if sp >= 0 else None, bool(both[i])))
    return out


# ─── blend ──────────────────────────────────────────────────────────────────────────────────
def blend_two_tier(fast: Sequence[tuple], quality: Sequence[tuple], blend_factor: float):
    """fast / quality: [(doc_id, index, score)] best-first.  Returns [(doc_id, index, score)]."""
    fb, fo = pack_ids([d for d, _, _ in fast])
    fi = _c([i for _, i, _ in fast], np.uint32)
    fs = _c([s for _, _, s in fast], np.float32)
    qb, qo = pack_ids([d for d, _, _ in quality])
    qi = _c([i for _, i, _ in quality], np.uint32)
    qs = _c([s for _, _, s in quality], np.float32)
    cap = max(1, len(fast) + len(quality))
    oi = np.empty(cap, dtype=np.uint32)
    os_ = np.empty(cap, dtype=np.float32)
    src = np.empty(cap, dtype=np.int64)
    n = lib().fso_blend_two_tier(_p(fb, _u8p), _p(fo, _u64p), _p(fi, _u32p), _p(fs, _f32p), len(fast),
                                 _p(qb, _u8p), _p(qo, _u64p), _p(qi, _u32p), _p(qs, _f32p), len(quality),
                                 blend_factor, _p(oi, _u32p), _p(os_, _f32p), _p(src, _i64p))
    out = []
    for i in range(n):
        s = int(src[i])
        doc = fast[s][0] if s >= 0 else quality[-s - 1][0]
        out.append((doc, int(oi[i]), np.float32(os_[i])))
    return out


def blend_two_tier_aligned(fast: Sequence[tuple], quality_scores: Sequence[Optional[float]],
                           blend_factor: float):
    fb, fo = pack_ids([d for d, _, _ in fast])
    fi = _c([i for _, i, _ in fast], np.uint32)
    fs = _c([s for _, _, s in fast], np.float32)
    present = _c([q is not None for q in quality_scores], np.uint8)
    qs = _c([0.0 if q is None else q for q in quality_scores], np.float32)
    cap = max(1, len(fast))
    oi = np.empty(cap, dtype=np.uint32)
    os_ = np.empty(cap, dtype=np.float32)
    src = np.empty(cap, dtype=np.int64)
    n = lib().fso_blend_two_tier_aligned(_p(fb, _u8p), _p(fo, _u64p), _p(fi, _u32p), _p(fs, _f32p),
                                         len(fast), _p(qs, _f32p), _p(present, _u8p), blend_factor,
                                         _p(oi, _u32p), _p(os_, _f32p), _p(src, _i64p))
    return [(fast[int(src[i])][0], int(oi[i]), np.float32(os_[i])) for i in range(n)]


# ─── potion ─────────────────────────────────────────────────────────────────────────────────
def potion_embed(table, ids) -> np.ndarray:
    table = _c(table, np.float32)
    vocab, dim = table.shape
    ids = _c(ids, np.uint32)
    out = np.empty(dim, dtype=np.float32)
    lib().fso_potion_embed(_p(table, _f32p), vocab, dim, _p(ids, _u32p), ids.size, _p(out, _f32p))
    return out


def l2_normalize(v) -> np.ndarray:
    v = _c(v, np.float32).copy()
    lib().fso_l2_normalize(_p(v, _f32p), v.size)
    return v


# ─── synthetic corpora (reference bench generators) ─────────────────────────────────────────
def raw_vector(seed: int, dim: int) -> np.ndarray:
    out = np.empty(dim, dtype=np.float32)
    lib().fso_raw_vector(seed, dim, _p(out, _f32p))
    return out


def normalize(v) -> np.ndarray:
    v = _c(v, np.float32).copy()
    lib().fso_normalize(_p(v, _f32p), v.size)
    return v


def synth_rows(kind: int, seed_base: int, row_start: int, n_rows: int, dim: int,
               n_centroids: int = 64, noise: float = 0.30, want_f32: bool = False,
               threads: Optional[int] = None):
    """kind 0 uniform / 1 clustered (fsvi_int8_two_pass.rs:199-231).  Returns (f16 bits, f32|None)."""
    f16 = np.empty((n_rows, dim), dtype=np.uint16)
    f32 = np.empty((n_rows, dim), dtype=np.float32) if want_f32 else None
    lib().fso_synth_rows(kind, seed_base, row_start, n_rows, dim, n_centroids, noise,
                         threads or host_threads(), _p(f32, _f32p), _p(f16, _u16p))
    return f16, f32


def clustered_query(q: int, dim: int, n_centroids: int = 64, noise: float = 0.30) -> np.ndarray:
    """fsvi_int8_two_pass.rs:285-287: make_vector(centroids, q % C, 0xdead_0000 + q)."""
    c = normalize(raw_vector(0xC0000000 + (q % n_centroids), dim))
    nz = raw_vector(0xDEAD0000 + q, dim)
    return normalize((c + np.float32(noise) * nz).astype(np.float32))
