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


# ─── header reader ──────────────────────────────────────────────────────────────────────────
def read_fsvi_header(path: str) -> dict:
    """The v1 header fields a host needs before opening a file on the GPU (lib.rs:4049-4144):
    dimension, quantization, compaction generation (reserved byte 0, lib.rs:4110), publication nonce
    (reserved bytes 1-2), record count, vectors offset.  CRC-checked."""
    with open(path, "rb") as f:
        data = f.read(4 + 2 + 2 + 65535 + 2 + 65535 + 4 + 1 + 3 + 8 + 8 + 4)
    if len(data) < 8 or data[:4] != FSVI_MAGIC:
        raise SearchError("IndexCorrupted", f"{path}: bad magic")
    (version,) = struct.unpack_from("<H", data, 4)
    if version == 2:  # identity-complete artifact (lib.rs:4229-4447): fixed 332-byte prefix, CRC at header_size - 4
        header_size, schema, quant, flags, nonce, dim, count, voff = struct.unpack_from("<IHBBHIQQ", data, 6)
        if header_size < 336:
            raise SearchError("IndexCorrupted", f"{path}: v2 header_size out of range")
        if len(data) < header_size:
            with open(path, "rb") as f:
                data = f.read(header_size)
        if len(data) < header_size:
            raise SearchError("IndexCorrupted", f"{path}: v2 header is truncated")
        (crc,) = struct.unpack_from("<I", data, header_size - 4)
        if zlib.crc32(data[:header_size - 4]) & 0xFFFFFFFF != crc:
            raise SearchError("IndexCorrupted", f"{path}: v2 header CRC mismatch")
        return dict(version=2, embedder_id="", embedder_revision="", dimension=dim, quantization=quant,
                    compaction_gen=0, publication_nonce=nonce, record_count=count, vectors_offset=voff)
    if version != FSVI_VERSION or len(data) < 8:
        raise SearchError("IndexCorrupted", f"{path}: unsupported FSVI version {version}")
    (n,) = struct.unpack_from("<H", data, 6)
    cur = 8
    embedder_id = data[cur:cur + n].decode("utf-8", "replace")
    cur += n
    (n,) = struct.unpack_from("<H", data, cur)
    cur += 2
    revision = data[cur:cur + n].decode("utf-8", "replace")
    cur += n
    dim, quant = struct.unpack_from("<IB", data, cur)
    reserved = data[cur + 5:cur + 8]
    count, voff = struct.unpack_from("<QQ", data, cur + 8)
    crc_end = cur + 24
    (crc,) = struct.unpack_from("<I", data, crc_end)
    if zlib.crc32(data[:crc_end]) & 0xFFFFFFFF != crc:
        raise SearchError("IndexCorrupted", f"{path}: header CRC mismatch")
    return dict(version=version, embedder_id=embedder_id, embedder_revision=revision, dimension=dim,
                quantization=quant, compaction_gen=reserved[0], publication_nonce=reserved[1] | (reserved[2] << 8),
                record_count=count, vectors_offset=voff)


# ─── WAL sidecar (crates/frankensearch-index/src/wal.rs) ────────────────────────────────────
WAL_MAGIC, WAL_BATCH_MAGIC, WAL_VERSION, WAL_HEADER_SIZE = b"FWAL", b"FWB1", 1, 20


def wal_path_for(fsvi_path: str) -> str:
    """wal.rs:575-579: `<index>.wal`."""
    return fsvi_path + ".wal"


def next_generation(current: int) -> int:
    """lib.rs:6156-6158."""
    return 1 if current == 255 else current + 1


def append_wal_batch(wal_path: str, entries, dimension: int, quantization: int = QUANT_F16,
                     compaction_gen: int = 0) -> None:
    """wal.rs:1183-1283 + write_wal_header :1294-1312: 20-byte header {magic, version u16, dimension u32,
    quantization u8, compaction_gen u8, reserved[4], crc32 of the first 16 bytes} on first use, then one
    batch {"FWB1", entry_count u32, entries {doc_id_len u16, doc_id, vector in the index quantization},
    crc32 of the batch bytes}."""
    import os

    fresh = not os.path.exists(wal_path) or os.path.getsize(wal_path) < WAL_HEADER_SIZE
    batch = bytearray(WAL_BATCH_MAGIC + struct.pack("<I", len(entries)))
    for doc_id, vec in entries:
        b = doc_id.encode("utf-8")
        if len(b) > 0xFFFF:
            raise SearchError("InvalidConfig", "doc_id byte length must fit in u16")
        v = np.ascontiguousarray(vec, dtype=np.float32).reshape(-1)
        if v.size != dimension:
            raise SearchError("DimensionMismatch", f"expected {dimension}, found {v.size}")
        batch += struct.pack("<H", len(b)) + b
        with np.errstate(over="ignore"):
            batch += v.astype(np.float16).tobytes() if quantization == QUANT_F16 else v.tobytes()
    batch += struct.pack("<I", zlib.crc32(bytes(batch)) & 0xFFFFFFFF)
    with open(wal_path, "wb" if fresh else "ab") as f:
        if fresh:
            head = WAL_MAGIC + struct.pack("<HIBB", WAL_VERSION, dimension, quantization, compaction_gen) + b"\0" * 4
            f.write(head + struct.pack("<I", zlib.crc32(head) & 0xFFFFFFFF))
        f.write(bytes(batch))


def read_wal(wal_path: str, dimension: int, quantization: int = QUANT_F16):
    """wal.rs:852-866 read_wal + :980-1130 parse: returns (entries [(doc_id, float32[dim])] in file order,
    compaction_gen, valid_len).  A missing or shorter-than-header file is no WAL; a bad header is an
    error (IndexCorrupted); a corrupt or truncated batch ends the replay there (crash tolerance)."""
    import os

    if not os.path.exists(wal_path):
        return [], 0, 0
    data = open(wal_path, "rb").read()
    if len(data) < WAL_HEADER_SIZE:
        return [], 0, 0

    def corrupt(why):
        return SearchError("IndexCorrupted", f"{wal_path}: {why}")

    if data[:4] != WAL_MAGIC:
        raise corrupt("bad magic bytes")
    version, dim, quant, gen = struct.unpack_from("<HIBB", data, 4)
    if version != WAL_VERSION:
        raise corrupt(f"version mismatch: expected {WAL_VERSION}, got {version}")
    if dim != dimension:
        raise corrupt(f"dimension mismatch: expected {dimension}, got {dim}")
    if quant not in (QUANT_F32, QUANT_F16):
        raise corrupt("unknown quantization")
    if quant != quantization:
        raise corrupt("quantization mismatch")
    if struct.unpack_from("<I", data, 16)[0] != zlib.crc32(data[:16]) & 0xFFFFFFFF:
        raise corrupt("header CRC mismatch")
    vector_bytes = dimension * (2 if quantization == QUANT_F16 else 4)
    entries, cur = [], WAL_HEADER_SIZE
    while cur + 8 <= len(data):
        if data[cur:cur + 4] != WAL_BATCH_MAGIC:
            break
        (count,) = struct.unpack_from("<I", data, cur + 4)
        c, batch, ok = cur + 8, [], True
        for _ in range(count):
            if c + 2 > len(data):
                ok = False
                break
            (n,) = struct.unpack_from("<H", data, c)
            c += 2
            if c + n + vector_bytes > len(data):
                ok = False
                break
            try:
                doc_id = data[c:c + n].decode("utf-8")
            except UnicodeDecodeError:
                ok = False
                break
            c += n
            raw = np.frombuffer(data, dtype=np.float16 if quantization == QUANT_F16 else np.float32,
                                count=dimension, offset=c)
            batch.append((doc_id, raw.astype(np.float32)))  # f16 -> f32 is exact (decode_vector, wal.rs:1132)
            c += vector_bytes
        if not ok or c + 4 > len(data) or struct.unpack_from("<I", data, c)[0] != zlib.crc32(data[cur:c]) & 0xFFFFFFFF:
            break
        entries.extend(batch)
        cur = c + 4
    return entries, gen, cur


def replay_wal_for(fsvi_path: str, header: Optional[dict] = None):
    """The WAL half of VectorIndex::open (lib.rs:1833-1878): entries of the sidecar with the LAST entry
    of each doc id kept (in file order of those last entries), or nothing when the sidecar is stale
    (its generation is not the successor of the main file's)."""
    h = header or read_fsvi_header(fsvi_path)
    entries, wal_gen, valid_len = read_wal(wal_path_for(fsvi_path), h["dimension"], h["quantization"])
    seen, kept = set(), []
    for doc_id, v in reversed(entries):
        if doc_id not in seen:
            seen.add(doc_id)
            kept.append((doc_id, v))
    kept.reverse()
    if valid_len > 0:
        stale = h["compaction_gen"] > 0 if wal_gen == 0 else wal_gen != next_generation(h["compaction_gen"])
        if stale:
            return []
    return kept
