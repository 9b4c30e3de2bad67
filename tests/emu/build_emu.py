"""TEST INFRASTRUCTURE: build the host-emulated copy of the engine's tree / rules device code.

`libagz_emu.so` compiles alphago.jl_b200/csrc/engine.cu with a plain host compiler and -DAGZ_EMU: the
warp-per-game functors run on 32 fibers (tests/emu/emu_runtime.cpp), "device memory" is host memory, and
the network / NCCL parts are compiled out.  It lets `pytest -m "not gpu"` exercise the exact device code
paths (Go rules, select/expand/backup, noise, pick, re-root, compaction) against the oracle where no GPU
exists.  It is never loaded by the product package, which has no CPU fallback.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "libagz_emu.so")


def build(force=False):
    srcs = [os.path.join(ROOT, "alphago.jl_b200", "csrc", f) for f in
            ("engine.cu", "ops.cuh", "tree.cuh", "go_bits.cuh", "go_rules.cuh", "rng.cuh", "simt.h", "devrt.h")]
    srcs += [os.path.join(HERE, "emu_runtime.cpp"), os.path.join(ROOT, "include", "agz.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in srcs):
        return OUT
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-DAGZ_EMU=1", "-ffp-contract=off", "-fno-fast-math",
           "-Wall", "-Wno-unused-function", "-Wno-unknown-pragmas",
           "-x", "c++", srcs[0], os.path.join(HERE, "emu_runtime.cpp"), "-o", OUT]
    subprocess.run(cmd, check=True, cwd=ROOT)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
