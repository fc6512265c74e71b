"""Host-side mirror of the reference query encoders over the C ABI.

`Model2VecEmbedder` mirrors crates/frankensearch-embed/src/model2vec_embedder.rs:67 (potion /
Model2Vec static embedder).  Tokenisation stays on the host (the reference uses the HF
`tokenizers` crate, model2vec_embedder.rs:288-294); gather + mean pool + L2 run on the GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import SearchError, check, ptr


class Model2VecEmbedder:
    """`tokenizer(text) -> list[int]` must encode WITHOUT special tokens
    (`encode_fast(text, false)`, model2vec_embedder.rs:288)."""

    def __init__(self, embeddings, tokenizer: Optional[Callable[[str], Sequence[int]]] = None, *,
                 device: int = 0, name: str = "potion-multilingual-128M"):
        table = np.ascontiguousarray(embeddings, dtype=np.float32)
        if table.ndim != 2 or table.size == 0:
            raise SearchError("InvalidConfig", "embeddings must be a non-empty [vocab, dim] matrix")
        self._vocab, self._dim = table.shape
        self._tokenizer = tokenizer
        self._name = name
        self._L = _ffi.lib()
        h = C.c_void_p()
        check(self._L.fsgpu_potion_create(ptr(table), self._vocab, self._dim, device, C.byref(h)))
        self._h = h

    def close(self) -> None:
        if self._h:
            self._L.fsgpu_potion_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # Embedder trait surface (crates/frankensearch-core/src/traits.rs:220-370)
    def dimension(self) -> int:
        return self._dim

    def id(self) -> str:
        return self._name

    def is_semantic(self) -> bool:
        return True

    def embed_token_ids_batch(self, batches: Sequence[Sequence[int]]) -> np.ndarray:
        """embed_token_ids for each id list (model2vec_embedder.rs:312-335); [B, dim] f32."""
        b = len(batches)
        if b == 0:
            return np.zeros((0, self._dim), dtype=np.float32)
        offsets = np.zeros(b + 1, dtype=np.uint64)
        offsets[1:] = np.cumsum([len(x) for x in batches], dtype=np.uint64)
        flat = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.uint32).reshape(-1) for x in batches])
                                    if offsets[-1] else np.zeros(1, dtype=np.uint32), dtype=np.uint32)
        out = np.zeros((b, self._dim), dtype=np.float32)
        check(self._L.fsgpu_potion_embed(self._h, ptr(flat), ptr(offsets), b, ptr(out)))
        return out

    def embed_token_ids(self, token_ids: Sequence[int]) -> np.ndarray:
        return self.embed_token_ids_batch([token_ids])[0]

    def embed_sync(self, text: str) -> np.ndarray:
        """embed_sync (model2vec_embedder.rs:280-310): empty text -> zero vector."""
        if text == "":
            return np.zeros(self._dim, dtype=np.float32)
        if self._tokenizer is None:
            raise SearchError("EmbeddingFailed", "no tokenizer configured")
        return self.embed_token_ids(list(self._tokenizer(text)))

    def embed_batch(self, texts: Sequence[str]) -> List[np.ndarray]:
        if self._tokenizer is None:
            raise SearchError("EmbeddingFailed", "no tokenizer configured")
        ids = [list(self._tokenizer(t)) if t != "" else [] for t in texts]
        return list(self.embed_token_ids_batch(ids))
