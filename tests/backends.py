"""Backend selection for the C-ABI tests.

  "cuda": the product library alphago.jl_b200/libagz.so on a B200 (tests marked `gpu`);
  "emu" : tests/emu/libagz_emu.so -- the SAME device code (tree / rules / features) compiled for the host fiber
          emulator, so the CPU suite exercises the kernels' logic against the oracle without a GPU.
"""
import pytest

import pkg

agz = pkg.load()


def _emu_lib():
    from emu.build_emu import build
    return build()


def lib_for(backend):
    if backend == "emu":
        return _emu_lib()
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return None  # default product library


BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]
