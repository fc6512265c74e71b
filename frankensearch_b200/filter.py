"""Host mirror of the reference's `SearchFilter` seam (crates/frankensearch-core/src/filter.rs).

Filters are evaluated per doc id on the host — except `BitsetFilter`, whose decision depends only on
the 8-byte doc-id hash and is evaluated on the device against the record-table hashes
(`fsgpu_search_top_k_hashes`), including the reference's selective gather arm
(crates/frankensearch-index/src/search.rs:1114-1255)."""
from __future__ import annotations

from typing import Callable, Iterable, Optional

from .types import fnv1a_hash


class BitsetFilter:
    """filter.rs:330-383: a document passes iff the FNV-1a hash of its doc id is in the set."""

    name = "bitset_filter"

    def __init__(self, hashes: Iterable[int]):
        self.hashes = frozenset(int(h) & 0xFFFFFFFFFFFFFFFF for h in hashes)

    @classmethod
    def from_hashes(cls, hashes: Iterable[int]) -> "BitsetFilter":
        return cls(hashes)

    @classmethod
    def from_doc_ids(cls, doc_ids: Iterable[str]) -> "BitsetFilter":
        return cls(fnv1a_hash(d.encode("utf-8")) for d in doc_ids)

    def matches(self, doc_id: str, metadata=None) -> bool:
        return fnv1a_hash(doc_id.encode("utf-8")) in self.hashes

    def matches_doc_id_hash(self, doc_id_hash: int, metadata=None) -> Optional[bool]:
        return doc_id_hash in self.hashes

    def candidate_hashes(self):
        return self.hashes

    def __call__(self, doc_id: str) -> bool:
        return self.matches(doc_id)


class PredicateFilter:
    """filter.rs PredicateFilter: an arbitrary predicate on the doc id (host-evaluated)."""

    def __init__(self, name: str, predicate: Callable[[str], bool]):
        self.name = name
        self._p = predicate

    def matches(self, doc_id: str, metadata=None) -> bool:
        return bool(self._p(doc_id))

    def matches_doc_id_hash(self, doc_id_hash: int, metadata=None) -> Optional[bool]:
        return None

    def candidate_hashes(self):
        return None

    def __call__(self, doc_id: str) -> bool:
        return self.matches(doc_id)
