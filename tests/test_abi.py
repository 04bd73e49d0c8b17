"""CPU-only checks of the drop-in boundary: libdawn_b200.so loads, exports every symbol that
include/dawn_index.h declares, refuses to work without a GPU (no CPU fallback), and the
product package never touches oracle/."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dawn_index.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dawn_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_reference_surface():
    names = declared_symbols()
    # one entry point per usearch::ffi call site in src/search/search_provider.rs
    for want in ("dawn_index_create", "dawn_index_free", "dawn_index_reserve", "dawn_index_add",
                 "dawn_index_add_batch", "dawn_index_search", "dawn_index_search_batch",
                 "dawn_index_size", "dawn_index_capacity", "dawn_index_dimensions",
                 "dawn_index_save", "dawn_index_load", "dawn_last_error"):
        assert want in names


def test_library_exports_every_declared_symbol(dawn):
    lib = dawn.load_library()
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/dawn_index.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", dawn.index.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (dawn_[a-z0-9_]+)", out))
    assert set(declared_symbols()) <= exported
    assert b"sm_100a" in lib.dawn_version()


def test_library_is_built_for_sm_100a_only(dawn):
    out = subprocess.run(["cuobjdump", "-lelf", dawn.index.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu(dawn):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is for hosts without one")
    with pytest.raises(dawn.DawnError) as ei:
        dawn.new_index(dawn.IndexOptions())
    assert ei.value.code == -2  # DAWN_ERR_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_argument_validation_needs_no_gpu(dawn):
    lib = dawn.load_library()
    h = C.c_void_p()
    opts = dawn.index._Options(128, 0, 0, 0, 0, 0, 0)  # wrong dimension
    assert lib.dawn_index_create(C.byref(opts), C.byref(h)) == -1
    assert b"384" in lib.dawn_last_error()
    opts = dawn.index._Options(384, 0, 7, 0, 0, 0, 0)  # unknown storage kind
    assert lib.dawn_index_create(C.byref(opts), C.byref(h)) == -1
    assert lib.dawn_index_size(None) == 0
    assert lib.dawn_index_reserve(None, 10) == -1


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "dawnsearch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                for line in text.splitlines():
                    code = line.split("//")[0].split("#")[0] if not f.endswith(".py") else line.split("#")[0]
                    assert not re.search(r"(import|from)\s+oracle|libdawn_oracle|dawn_oracle_\w+\s*\(|#include\s+\"[^\"]*oracle", code), (f, line)
