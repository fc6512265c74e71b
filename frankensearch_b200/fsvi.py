"""FSVI v1 container: writer + header reader (host-side file format plumbing).

Layout and ordering follow the reference writer (crates/frankensearch-index/src/lib.rs:6-43,
:3752-3762 stable sort by (fnv1a(doc_id), doc_id), :3914-3943 section order, :5714-5768 header,
:6114 CRC-32).  The scan itself never touches this module: `GpuVectorIndex.open` parses the file
inside libfsgpu.so and uploads the slab.
"""
from __future__ import annotations

import struct
import zlib
from typing import Optional, Sequence

import numpy as np

from ._ffi import SearchError
from .types import fnv1a_hash

FSVI_MAGIC = b"FSVI"
FSVI_VERSION = 1
QUANT_F32, QUANT_F16 = 0, 1


def write_fsvi_v1(path: str, embedder_id: str, dimension: int, doc_ids: Sequence[str], vectors,
                  *, embedder_revision: str = "", tombstones: Optional[Sequence[bool]] = None,
                  quantization: int = QUANT_F16) -> np.ndarray:
    """VectorIndex::create + write_record* + finish.  Returns the permutation applied by the
    writer's sort (file row -> input position)."""
    v = np.ascontiguousarray(vectors, dtype=np.float32)
    if v.ndim != 2 or v.shape[1] != dimension:
        raise SearchError("DimensionMismatch", f"expected {dimension}, found {v.shape[-1] if v.ndim else 0}")
    if len(doc_ids) != v.shape[0]:
        raise SearchError("InvalidConfig", "doc_ids and vectors differ in length")
    if not embedder_id:
        raise SearchError("InvalidConfig", "embedder_id cannot be empty")
    if v.size and not np.isfinite(v).all():  # lib.rs:3647-3653
        raise SearchError("InvalidConfig", "all embedding values must be finite")
    if v.shape[0] and (np.square(v, dtype=np.float32).sum(axis=1) <= 0).any():  # lib.rs:3654-3660
        raise SearchError("InvalidConfig", "embedding norm must be non-zero and finite")
    ids_b = [d.encode("utf-8") for d in doc_ids]
    if any(len(b) > 0xFFFF for b in ids_b):
        raise SearchError("InvalidConfig", "doc_id byte length must fit in u16")
    hashes = [fnv1a_hash(b) for b in ids_b]
    order = sorted(range(len(ids_b)), key=lambda i: (hashes[i], ids_b[i]))  # stable (lib.rs:3758-3762)
    flags = [0] * len(ids_b) if tombstones is None else [1 if t else 0 for t in tombstones]

    eid, erev = embedder_id.encode(), embedder_revision.encode()
    header_len = 4 + 2 + 2 + len(eid) + 2 + len(erev) + 4 + 1 + 3 + 8 + 8 + 4
    strings = b"".join(ids_b[i] for i in order)
    unpadded = header_len + 16 * len(order) + len(strings)
    vectors_offset = (unpadded + 63) // 64 * 64
    head = (FSVI_MAGIC + struct.pack("<HH", FSVI_VERSION, len(eid)) + eid + struct.pack("<H", len(erev)) + erev +
            struct.pack("<IB", dimension, quantization) + b"\0\0\0" + struct.pack("<QQ", len(order), vectors_offset))
    head += struct.pack("<I", zlib.crc32(head) & 0xFFFFFFFF)
    records = bytearray()
    off = 0
    for i in order:
        records += struct.pack("<QIHH", hashes[i], off, len(ids_b[i]), flags[i])
        off += len(ids_b[i])
    perm = np.asarray(order, dtype=np.int64)
    body = v[perm] if len(order) else v
    with np.errstate(over="ignore"):
        slab = body.astype(np.float16).tobytes() if quantization == QUANT_F16 else body.tobytes()
    with open(path, "wb") as f:
        f.write(head)
        f.write(records)
        f.write(strings)
        f.write(b"\0" * (vectors_offset - unpadded))
        f.write(slab)
    return perm
