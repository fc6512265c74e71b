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


class MiniLmEmbedder:
    """Mirror of `FastEmbedEmbedder` for all-MiniLM-L6-v2
    (crates/frankensearch-embed/src/fastembed_embedder.rs:169, :317-398): BERT forward, attention-mask
    mean pool, L2, adapter L2 — all in libfsgpu.so (tcgen05 GEMMs, minilm_kernels.cuh).

    `weights` maps Hugging Face `BertModel` state-dict names to float32 arrays
    (`embeddings.word_embeddings.weight`, `encoder.layer.0.attention.self.query.weight`, ...), e.g.
    `{k: v.numpy() for k, v in BertModel.state_dict().items()}` or a loaded `model.safetensors`.
    `tokenizer(text) -> list[int]` must return the full id sequence including [CLS] / [SEP],
    truncated to 512 (model_manifest.rs:74-80); tokenisation stays on the host."""

    DIM = 384

    def __init__(self, weights, tokenizer: Optional[Callable[[str], Sequence[int]]] = None, *, device: int = 0,
                 heads: int = 12, ln_eps: float = 1e-12, name: str = "all-MiniLM-L6-v2"):
        def get(key):
            for prefix in ("", "bert.", "0.auto_model."):
                if prefix + key in weights:
                    return np.ascontiguousarray(np.asarray(weights[prefix + key], dtype=np.float32))
            raise SearchError("InvalidConfig", f"minilm: weight {key!r} is missing")

        self._L = _ffi.lib()
        self._tokenizer = tokenizer
        self._name = name
        keep = []  # host arrays must outlive the create call

        def ptr_of(a):
            keep.append(a)
            return a.ctypes.data

        word = get("embeddings.word_embeddings.weight")
        n_layers = 0
        while any(k.endswith(f"encoder.layer.{n_layers}.attention.self.query.weight") for k in weights):
            n_layers += 1
        if n_layers == 0:
            raise SearchError("InvalidConfig", "minilm: no encoder layers in the weight map")
        layers = (_ffi.MiniLmLayerWeights * n_layers)()
        inter = 0
        for i in range(n_layers):
            p = f"encoder.layer.{i}."
            q, k, v = (get(p + f"attention.self.{n}.weight") for n in ("query", "key", "value"))
            qb, kb, vb = (get(p + f"attention.self.{n}.bias") for n in ("query", "key", "value"))
            L = layers[i]
            L.qkv_w = ptr_of(np.ascontiguousarray(np.concatenate([q, k, v], axis=0)))
            L.qkv_b = ptr_of(np.ascontiguousarray(np.concatenate([qb, kb, vb])))
            L.attn_out_w = ptr_of(get(p + "attention.output.dense.weight"))
            L.attn_out_b = ptr_of(get(p + "attention.output.dense.bias"))
            L.attn_ln_g = ptr_of(get(p + "attention.output.LayerNorm.weight"))
            L.attn_ln_b = ptr_of(get(p + "attention.output.LayerNorm.bias"))
            w_in = get(p + "intermediate.dense.weight")
            inter = w_in.shape[0]
            L.ffn_in_w = ptr_of(w_in)
            L.ffn_in_b = ptr_of(get(p + "intermediate.dense.bias"))
            L.ffn_out_w = ptr_of(get(p + "output.dense.weight"))
            L.ffn_out_b = ptr_of(get(p + "output.dense.bias"))
            L.ffn_ln_g = ptr_of(get(p + "output.LayerNorm.weight"))
            L.ffn_ln_b = ptr_of(get(p + "output.LayerNorm.bias"))
        pos = get("embeddings.position_embeddings.weight")
        w = _ffi.MiniLmWeights()
        w.vocab_size, w.hidden = word.shape
        w.max_positions = pos.shape[0]
        w.n_layers, w.heads, w.intermediate, w.ln_eps = n_layers, heads, inter, ln_eps
        w.word_emb, w.pos_emb = ptr_of(word), ptr_of(pos)
        w.type_emb = ptr_of(get("embeddings.token_type_embeddings.weight"))
        w.emb_ln_g = ptr_of(get("embeddings.LayerNorm.weight"))
        w.emb_ln_b = ptr_of(get("embeddings.LayerNorm.bias"))
        w.layers = layers
        self._max_positions = int(pos.shape[0])
        h = C.c_void_p()
        check(self._L.fsgpu_minilm_create(C.byref(w), device, C.byref(h)))
        self._h = h

    @classmethod
    def from_safetensors(cls, path: str, tokenizer: Optional[Callable[[str], Sequence[int]]] = None, *, device: int = 0,
                         max_positions: int = 512, name: str = "all-MiniLM-L6-v2") -> "MiniLmEmbedder":
        """Weights straight from `model.safetensors` through the C ABI (fsgpu_minilm_load): the file the
        reference pins for all-MiniLM-L6-v2 (model_manifest.rs:343-349)."""
        self = cls.__new__(cls)
        self._L = _ffi.lib()
        self._tokenizer = tokenizer
        self._name = name
        self._max_positions = int(max_positions)
        h = C.c_void_p()
        check(self._L.fsgpu_minilm_load(path.encode(), device, C.byref(h)))
        self._h = h
        return self

    @classmethod
    def from_model_dir(cls, model_dir: str, *, device: int = 0) -> "MiniLmEmbedder":
        """A model directory as the reference downloads it (`model.safetensors` + `tokenizer.json`,
        model_manifest.rs:327-349): weights through fsgpu_minilm_load, tokenisation by the Hugging Face
        `tokenizers` library under the adapter's sequence policy (minilm_token_ids)."""
        import os

        from tokenizers import Tokenizer

        tok = Tokenizer.from_file(os.path.join(model_dir, "tokenizer.json"))
        return cls.from_safetensors(os.path.join(model_dir, "model.safetensors"),
                                    tokenizer=lambda text: minilm_token_ids(tok, text), device=device)

    def close(self) -> None:
        if self._h:
            self._L.fsgpu_minilm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # Embedder trait surface (crates/frankensearch-core/src/traits.rs:220-370)
    def dimension(self) -> int:
        return self.DIM

    def id(self) -> str:
        return self._name

    def is_semantic(self) -> bool:
        return True

    def embed_token_ids_batch(self, batches: Sequence[Sequence[int]]) -> np.ndarray:
        """Batch-longest padding (model_manifest.rs:74-80); an empty id list gives the zero vector
        (fastembed_embedder.rs:432-434).  Returns [B, 384] float32, L2-normalised."""
        b = len(batches)
        if b == 0:
            return np.zeros((0, self.DIM), dtype=np.float32)
        lens = np.array([min(len(x), self._max_positions) for x in batches], dtype=np.int32)
        max_len = max(int(lens.max()), 1)
        ids = np.zeros((b, max_len), dtype=np.int32)
        for i, x in enumerate(batches):
            ids[i, :lens[i]] = np.asarray(x[:lens[i]], dtype=np.int32)
        out = np.zeros((b, self.DIM), dtype=np.float32)
        check(self._L.fsgpu_minilm_embed(self._h, ptr(ids), ptr(lens), b, max_len, ptr(out)))
        return out

    def embed_token_ids(self, token_ids: Sequence[int]) -> np.ndarray:
        return self.embed_token_ids_batch([token_ids])[0]

    def embed_device(self, d_ids, d_lens, stream=None):
        """Device-resident form: int32 CUDA tensors ids [B, T] and lens [B] -> float32 [B, 384]."""
        import torch

        b, t = d_ids.shape
        out = torch.empty((b, self.DIM), dtype=torch.float32, device=d_ids.device)
        s = torch.cuda.current_stream(d_ids.device).cuda_stream if stream is None else stream
        check(self._L.fsgpu_minilm_embed_device(self._h, d_ids.data_ptr(), d_lens.data_ptr(), b, t, out.data_ptr(), s))
        return out

    def embed_sync(self, text: str) -> np.ndarray:
        if text == "":
            return np.zeros(self.DIM, dtype=np.float32)
        if self._tokenizer is None:
            raise SearchError("EmbeddingFailed", "no tokenizer configured")
        return self.embed_token_ids(list(self._tokenizer(text)))

    def embed_batch(self, texts: Sequence[str]) -> List[np.ndarray]:
        if self._tokenizer is None:
            raise SearchError("EmbeddingFailed", "no tokenizer configured")
        return list(self.embed_token_ids_batch([list(self._tokenizer(t)) if t != "" else [] for t in texts]))

    def profile_enable(self, on: bool = True) -> None:
        check(self._L.fsgpu_minilm_profile_enable(self._h, 1 if on else 0))

    def profile_read(self, reset: bool = True) -> dict:
        p = _ffi.MiniLmProfile()
        check(self._L.fsgpu_minilm_profile_read(self._h, C.byref(p), 1 if reset else 0))
        return dict(gemm_launches=int(p.gemm_launches), other_launches=int(p.other_launches),
                    gemm_flops=float(p.gemm_flops), gemm_ms=float(p.gemm_ms))


# ─── the host half of the MiniLM input contract ───────────────────────────────────────────────
MINILM_MAX_LENGTH = 512  # FASTEMBED_MAX_LENGTH_V1 (crates/frankensearch-embed/src/model_manifest.rs:74)
MINILM_CLS, MINILM_SEP = 101, 102  # [CLS] / [SEP] of the bert-base-uncased vocabulary the model ships


def minilm_sequence(wordpiece_ids: Sequence[int], *, cls_id: int = MINILM_CLS, sep_id: int = MINILM_SEP,
                    max_length: int = MINILM_MAX_LENGTH) -> List[int]:
    """`bert-tokenizer-special-tokens=true` + `max-length=512;longest-first` (model_manifest.rs:74-80,
    :300-304): `[CLS] w_1 .. w_n [SEP]`, truncated so that the whole sequence — specials included — fits
    `max_length` (a single sequence loses its LAST word pieces, the specials stay)."""
    room = max(max_length - 2, 0)
    return [cls_id] + [int(t) for t in wordpiece_ids[:room]] + [sep_id]


def minilm_token_ids(tokenizer, text: str) -> List[int]:
    """Token ids of `text` for MiniLmEmbedder under the adapter's policy, from a Hugging Face
    `tokenizers.Tokenizer` (the reference's tokenizer family: huggingface-tokenizers-json-v1).  The empty
    string is answered before tokenisation (zero vector, fastembed_embedder.rs:432-434)."""
    if text == "":
        return []
    enc = tokenizer.encode(text, add_special_tokens=False)
    cls_id = tokenizer.token_to_id("[CLS]")
    sep_id = tokenizer.token_to_id("[SEP]")
    return minilm_sequence(enc.ids, cls_id=MINILM_CLS if cls_id is None else cls_id,
                           sep_id=MINILM_SEP if sep_id is None else sep_id)


def minilm_pad_batch(sequences: Sequence[Sequence[int]], pad_id: int = 0):
    """`batch-longest-padding`: int32 ids [B, longest] padded with `pad_id`, int32 lens [B]; the attention
    mask is `position < len` (fsgpu_minilm_embed ignores slots at or past len)."""
    lens = np.array([len(s) for s in sequences], dtype=np.int32)
    t = max(int(lens.max()) if len(sequences) else 0, 1)
    ids = np.full((len(sequences), t), pad_id, dtype=np.int32)
    for i, srow in enumerate(sequences):
        ids[i, :len(srow)] = np.asarray(srow, dtype=np.int32)
    return ids, lens
