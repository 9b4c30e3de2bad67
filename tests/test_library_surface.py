"""CPU checks of the boundary: libagz.so loads without a GPU, exports every symbol include/agz.h declares,
refuses to create an engine without CUDA (no CPU fallback), and the ctypes structs match the header."""
import ctypes
import os
import re

import pytest

import pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
agz = pkg.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "agz.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(agz_[a-z0-9_]+)\s*\(", src)))


def test_binding_lists_every_header_symbol():
    assert sorted(agz.SYMBOLS) == header_symbols()


def test_product_library_exports_every_symbol():
    if not os.path.exists(agz.LIB_PATH):
        pytest.skip("libagz.so not built (run python __graft_entry__.py)")
    lib = ctypes.CDLL(agz.LIB_PATH)
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert lib.agz_version() >= 100


def test_emulator_library_exports_every_symbol():
    from emu.build_emu import build
    lib = ctypes.CDLL(build())
    for name in header_symbols():
        assert hasattr(lib, name), name


def test_no_cpu_fallback():
    """Without a CUDA device the product library must fail loudly instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    if not os.path.exists(agz.LIB_PATH):
        pytest.skip("libagz.so not built")
    with pytest.raises(agz.AgzError) as ei:
        agz.Engine(9, n_games=1)
    assert ei.value.code == 3 and "no CPU fallback" in str(ei.value)


def test_struct_sizes_match_header():
    # sizes computed from the C declarations in include/agz.h
    assert ctypes.sizeof(agz.Config) == 136
    assert ctypes.sizeof(agz.Position) == 2924
    assert ctypes.sizeof(agz.NodeView) == 9456
    assert ctypes.sizeof(agz.binding.GameHeader) == 32
    assert ctypes.sizeof(agz.binding.Progress) == 64


def test_product_package_never_imports_the_oracle():
    pkgdir = os.path.join(ROOT, "alphago.jl_b200")
    for dirpath, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
