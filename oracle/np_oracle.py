"""NumPy mirror of the C++ oracle's arithmetic — an independent second statement used only to
cross-check oracle/fs_oracle.cpp in tests/ (TEST INFRASTRUCTURE ONLY, never imported by the product).

Every float32 operation below is a single correctly-rounded IEEE op (NumPy never fuses
multiply-add), so the accumulation tree is the reference's, bit for bit
(crates/frankensearch-index/src/simd.rs:398-446).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def decode_f16(bits: np.ndarray) -> np.ndarray:
    return np.asarray(bits, dtype=np.uint16).view(np.float16).astype(np.float32)


def encode_f16(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        return np.asarray(x, dtype=np.float32).astype(np.float16).view(np.uint16)


def _reduce8(v: np.ndarray, order: int) -> np.float32:
    v = [F32(x) for x in v]
    if order == 1:
        return F32(F32(F32(v[0] + v[4]) + F32(v[2] + v[6])) + F32(F32(v[1] + v[5]) + F32(v[3] + v[7])))
    if order == 2:
        return F32(F32(F32(F32(v[0] + v[1]) + v[2]) + v[3]) + F32(F32(F32(v[4] + v[5]) + v[6]) + v[7]))
    if order == 3:
        return F32(F32(F32(v[0] + v[2]) + F32(v[1] + v[3])) + F32(F32(v[4] + v[6]) + F32(v[5] + v[7])))
    if order == 4:
        acc = v[0]
        for x in v[1:]:
            acc = F32(acc + x)
        return acc
    return F32(F32(F32(v[0] + v[1]) + F32(v[2] + v[3])) + F32(F32(v[4] + v[5]) + F32(v[6] + v[7])))


def dot_f16_f32(row_bits: np.ndarray, query: np.ndarray, reduce_order: int = 0,
                tail_fma: bool = True) -> np.float32:
    """One row.  Slow (Python loop over chunks) — small cases only."""
    x = decode_f16(row_bits)
    q = np.asarray(query, dtype=np.float32)
    dim = q.size
    chunks = dim // 8
    s = np.zeros((4, 8), dtype=np.float32)
    with np.errstate(all="ignore"):
        c = 0
        while c + 4 <= chunks:
            for a in range(4):
                lo = (c + a) * 8
                s[a] = s[a] + x[lo:lo + 8] * q[lo:lo + 8]
            c += 4
        while c < chunks:
            lo = c * 8
            s[0] = s[0] + x[lo:lo + 8] * q[lo:lo + 8]
            c += 1
        v = (s[0] + s[1]) + (s[2] + s[3])
        result = _reduce8(v, reduce_order)
        for e in range(chunks * 8, dim):
            if tail_fma:
                # fused multiply-add == exact product in f64 (24x11-bit significands fit) + one
                # rounding of the f64 sum to f32 is NOT always the same as fmaf; use exact integers.
                result = _fma_f32(x[e], q[e], result)
            else:
                result = F32(result + F32(x[e] * q[e]))
    return F32(result)


def _fma_f32(a, b, c) -> np.float32:
    """Correctly rounded a*b+c for float32 via exact rational arithmetic."""
    from fractions import Fraction
    a, b, c = float(a), float(b), float(c)
    if not (np.isfinite(a) and np.isfinite(b) and np.isfinite(c)):
        return F32(a * b + c)
    exact = Fraction(a) * Fraction(b) + Fraction(c)
    if exact == 0:
        return F32(a * b + c)  # signed-zero rules of the plain ops agree with fma here
    # round exact rational to nearest-even float32
    f = float(exact)  # correctly rounded to f64 (Fraction.__float__ is exact-rounding)
    r = F32(f)
    # repair double rounding: compare neighbours
    lo, hi = np.nextafter(r, F32(-np.inf)), np.nextafter(r, F32(np.inf))
    best = min((r, lo, hi), key=lambda t: (abs(Fraction(float(t)) - exact), int(t.view(np.uint32)) & 1))
    return F32(best)


def dot_rows(slab_bits: np.ndarray, query: np.ndarray, reduce_order: int = 0) -> np.ndarray:
    """Vectorised over rows for dim % 32 == 0 (the no-tail case): scores[n] float32."""
    x = decode_f16(slab_bits)
    n, dim = x.shape
    assert dim % 32 == 0
    q = np.asarray(query, dtype=np.float32)
    with np.errstate(all="ignore"):
        p = (x * q[None, :]).reshape(n, dim // 32, 4, 8)       # [n, j, a, l]
        s = np.zeros((n, 4, 8), dtype=np.float32)
        for j in range(dim // 32):
            s = s + p[:, j]
        v = (s[:, 0] + s[:, 1]) + (s[:, 2] + s[:, 3])            # [n, 8]
        if reduce_order == 1:
            return ((v[:, 0] + v[:, 4]) + (v[:, 2] + v[:, 6])) + ((v[:, 1] + v[:, 5]) + (v[:, 3] + v[:, 7]))
        if reduce_order == 2:
            return (((v[:, 0] + v[:, 1]) + v[:, 2]) + v[:, 3]) + (((v[:, 4] + v[:, 5]) + v[:, 6]) + v[:, 7])
        if reduce_order == 3:
            return ((v[:, 0] + v[:, 2]) + (v[:, 1] + v[:, 3])) + ((v[:, 4] + v[:, 6]) + (v[:, 5] + v[:, 7]))
        if reduce_order == 4:
            acc = v[:, 0]
            for l in range(1, 8):
                acc = acc + v[:, l]
            return acc
        return ((v[:, 0] + v[:, 1]) + (v[:, 2] + v[:, 3])) + ((v[:, 4] + v[:, 5]) + (v[:, 6] + v[:, 7]))


def order_keys(scores: np.ndarray, rows: np.ndarray) -> np.ndarray:
    """Ascending-u64 sort key == best-first order of search.rs:1673-1678."""
    s = np.asarray(scores, dtype=np.float32).copy()
    s[np.isnan(s)] = -np.inf
    u = s.view(np.uint32).astype(np.uint64)
    neg = (u >> np.uint64(31)) != 0
    asc = np.where(neg, u ^ np.uint64(0xFFFFFFFF), u ^ np.uint64(0x80000000)) & np.uint64(0xFFFFFFFF)
    desc = (~asc) & np.uint64(0xFFFFFFFF)
    return (desc << np.uint64(32)) | np.asarray(rows, dtype=np.uint64)


def top_k(scores: np.ndarray, limit: int, live: np.ndarray | None = None):
    rows = np.arange(scores.size, dtype=np.uint64)
    if live is not None:
        rows = rows[live]
        scores = scores[live]
    keys = order_keys(scores, rows)
    idx = np.argsort(keys, kind="stable")[:limit]
    return rows[idx], np.asarray(scores, dtype=np.float32)[idx]


def dot_f32_f32(a: np.ndarray, b: np.ndarray, reduce_order: int = 0) -> np.float32:
    """dot_product_f32_f32 (simd.rs:161-222, :1559-1587): left-over 8-chunks are added after the
    four accumulators are combined; scalar tail is mul then add."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    dim = a.size
    chunks, groups = dim // 8, dim // 32
    s = np.zeros((4, 8), dtype=np.float32)
    with np.errstate(all="ignore"):
        for g in range(groups):
            for acc in range(4):
                lo = g * 32 + acc * 8
                s[acc] = s[acc] + a[lo:lo + 8] * b[lo:lo + 8]
        v = (s[0] + s[1]) + (s[2] + s[3])
        for c in range(groups * 4, chunks):
            v = v + a[c * 8:c * 8 + 8] * b[c * 8:c * 8 + 8]
        result = _reduce8(v, reduce_order)
        for e in range(chunks * 8, dim):
            result = F32(result + F32(a[e] * b[e]))
    return F32(result)


def dot_f32_bytes_f32(a: np.ndarray, b: np.ndarray, reduce_order: int = 0) -> np.float32:
    """dot_product_f32_bytes_f32 (simd.rs:581-760), the score of a row of an f32-quantised FSVI slab
    (search.rs:1300-1321): the tree of dot_f32_f32 with a `mul_add` scalar tail."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    dim = a.size
    chunks, groups = dim // 8, dim // 32
    s = np.zeros((4, 8), dtype=np.float32)
    with np.errstate(all="ignore"):
        for g in range(groups):
            for acc in range(4):
                lo = g * 32 + acc * 8
                s[acc] = s[acc] + a[lo:lo + 8] * b[lo:lo + 8]
        v = (s[0] + s[1]) + (s[2] + s[3])
        for c in range(groups * 4, chunks):
            v = v + a[c * 8:c * 8 + 8] * b[c * 8:c * 8 + 8]
        result = _reduce8(v, reduce_order)
        for e in range(chunks * 8, dim):
            result = _fma_f32(a[e], b[e], result)
    return F32(result)


def resolve_sorted_entries(rows, scores, main_doc_ids, wal_doc_ids, tombstones=None):
    """The doc-id half of resolve_sorted_entries (search.rs:1503-1558) over best-first winners:
    a WAL winner (row >= len(main_doc_ids)) is dropped if its doc id was already emitted; a main
    winner is dropped if tombstoned, if ANY resident WAL row carries its doc id (shadowing), or if
    its doc id was already emitted.  Returns [(row, score, doc_id)]."""
    n = len(main_doc_ids)
    wal_set = set(wal_doc_ids)
    seen, out = set(), []
    for r, sc in zip(rows, scores):
        r = int(r)
        if r >= n:
            doc = wal_doc_ids[r - n]
            if doc in seen:
                continue
            seen.add(doc)
        else:
            if tombstones is not None and tombstones[r]:
                continue
            doc = main_doc_ids[r]
            if doc in wal_set or doc in seen:
                continue
            seen.add(doc)
        out.append((r, np.float32(sc), doc))
    return out


def quantize_f16_slab_to_i8(slab_bits: np.ndarray):
    """quantize_f16_slab_to_i8_generic (simd.rs:1842-1859): one corpus-wide scale 127 / max|x|,
    code = clamp(round_half_away_from_zero(x * scale), -127, 127).  Returns (codes int8, max_abs f32)."""
    x = decode_f16(slab_bits)
    max_abs = F32(np.max(np.abs(x))) if x.size else F32(0.0)
    if not max_abs > 0:
        return np.zeros(x.shape, dtype=np.int8), F32(0.0)
    scale = F32(F32(127.0) / max_abs)
    y = (x * scale).astype(np.float32)            # one f32 rounding, as `x.to_f32() * scale`
    # f32::round = half away from zero; the +0.5 is done in f64 so it cannot itself round up
    r = np.sign(y) * np.floor(np.abs(y).astype(np.float64) + 0.5)
    return np.clip(r, -127.0, 127.0).astype(np.int8), max_abs
