"""CPU: the C-ABI library loads and exports every symbol include/fsgpu.h declares; entry points
fail loudly (SubsystemError) without a GPU — there is no CPU fallback."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fsgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fsgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from frankensearch_b200 import _ffi

    L = _ffi.lib()
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/fsgpu.h but not exported by libfsgpu.so"
    assert sorted(_ffi.EXPORTS) == syms, "frankensearch_b200/_ffi.py EXPORTS is out of date"
    assert L.fsgpu_abi_version() == _ffi.ABI_VERSION == 4


def test_struct_layouts_match_header():
    from frankensearch_b200 import _ffi

    assert C.sizeof(_ffi.Hit) == 8
    assert C.sizeof(_ffi.FusedHitC) == 32
    assert C.sizeof(_ffi.IndexOptions) == 32
    assert C.sizeof(_ffi.RrfConfigC) == 32


def test_no_cpu_fallback_without_gpu(cuda_ok):
    if cuda_ok:
        pytest.skip("a GPU is present; the loud-failure path is exercised on the CPU box")
    import frankensearch_b200 as fs

    with pytest.raises(fs.SearchError) as e:
        fs.GpuVectorIndex.from_vectors(["a"], np.ones((1, 8), dtype=np.float32))
    assert e.value.kind == "SubsystemError"
    with pytest.raises(fs.SearchError):
        fs.rrf_fuse([fs.ScoredResult("a", 1.0)], [], 10)
    with pytest.raises(fs.SearchError):
        fs.Model2VecEmbedder(np.ones((4, 8), dtype=np.float32))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under frankensearch_b200/ may reference it."""
    pkg = os.path.join(ROOT, "frankensearch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "fs_oracle" not in text and "np_oracle" not in text and "libfs_oracle" not in text, f
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_header_is_plain_c_and_the_c_example_compiles():
    """include/fsgpu.h must be consumable by a C99 compiler (the boundary a Rust/C host binds);
    examples/c_abi_smoke.c is the plain-C call sequence."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = os.path.join(ROOT, "examples", "c_abi_smoke.c")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_rust_binding_source_lists_every_hot_path_symbol():
    """bindings/rust (unbuilt here: no cargo) must stay in step with the header for the entry points
    INTEGRATION.md wires up."""
    text = open(os.path.join(ROOT, "bindings", "rust", "src", "lib.rs")).read()
    bound = set(re.findall(r"pub fn (fsgpu_[a-z0-9_]+)\s*\(", text))
    assert bound <= set(declared_symbols()), bound - set(declared_symbols())
    for must in ("fsgpu_search_top_k", "fsgpu_search_top_k_filtered", "fsgpu_scores_for_rows", "fsgpu_rrf_fuse",
                 "fsgpu_blend_two_tier", "fsgpu_potion_embed", "fsgpu_minilm_embed", "fsgpu_index_open_fsvi"):
        assert must in bound, must


def test_graft_entry_build_passes():
    """`__graft_entry__.build()` is what the driver runs on the CPU box every round: the (incremental)
    nvcc + gcc build must succeed and its own ABI check must agree with the library."""
    import importlib
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    entry = importlib.import_module("__graft_entry__")
    entry.build()
    from frankensearch_b200 import _ffi

    assert _ffi.lib().fsgpu_abi_version() == _ffi.ABI_VERSION
