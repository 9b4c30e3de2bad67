"""The PODs that cross the C ABI must have ONE layout in all three places they are written down: include/agz.h (compiled here with
gcc into a tiny program that prints offsetof / sizeof of every field), the ctypes mirror (alphago.jl_b200/binding.py) and the Julia
mirror (julia/AlphaGoB200.jl, parsed: field order and types).  A size-only check would miss two swapped fields of equal width."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import pkg  # noqa: E402

agz = pkg.load()
B = agz.binding

STRUCTS = {"agz_config": B.Config, "agz_position": B.Position, "agz_node_view": B.NodeView, "agz_game_header": B.GameHeader,
           "agz_progress": B.Progress}


def header_fields():
    """{struct: [(field, c_type, array_suffix)]} parsed from include/agz.h."""
    src = open(os.path.join(ROOT, "include", "agz.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"typedef struct (\w+) \{(.*?)\} \1;", src, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            ty, rest = decl.split(None, 1)
            for name in rest.split(","):
                name = name.strip()
                base = re.match(r"\w+", name).group(0)
                fields.append((base, ty, name[len(base):]))
        out[m.group(1)] = fields
    return out


def c_layout(fields):
    prog = ['#include <stddef.h>', '#include <stdio.h>', '#include "agz.h"', 'int main(void) {']
    for s, fl in fields.items():
        prog.append('  printf("%s . %%zu %%zu\\n", (size_t)0, sizeof(%s));' % (s, s))
        for name, _, _ in fl:
            prog.append('  printf("%s %s %%zu %%zu\\n", offsetof(%s, %s), sizeof(((%s*)0)->%s));' % (s, name, s, name, s, name))
    prog += ['  return 0;', '}']
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "layout.c"), os.path.join(d, "layout")
        open(src, "w").write("\n".join(prog))
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        lines = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split("\n")
    lay = {}
    for ln in lines:
        if ln.strip():
            s, f, off, size = ln.split()
            lay[(s, f)] = (int(off), int(size))
    return lay


def test_header_is_plain_c_and_ctypes_offsets_match():
    fields = header_fields()
    assert set(STRUCTS) <= set(fields), "a struct of agz.h is missing from the test"
    lay = c_layout(fields)                                            # also proves agz.h compiles as C99 with -Wall -Werror
    for s, ct in STRUCTS.items():
        assert C.sizeof(ct) == lay[(s, ".")][1], s
        assert [f[0] for f in fields[s]] == [f[0] for f in ct._fields_], "%s: field order differs between agz.h and binding.py" % s
        for name, _, _ in fields[s]:
            d = getattr(ct, name)
            assert (d.offset, d.size) == lay[(s, name)], "%s.%s: ctypes (%d, %d) vs C %s" % (s, name, d.offset, d.size, lay[(s, name)])


JULIA_TYPES = {"int32_t": "Int32", "int64_t": "Int64", "uint64_t": "UInt64", "float": "Float32", "double": "Float64"}


def test_julia_struct_mirrors_match_header():
    """julia/AlphaGoB200.jl cannot be run here; its isbits structs must at least list the header's fields in order with the matching
    Julia types (Julia lays isbits structs out like C)."""
    fields = header_fields()
    jl = open(os.path.join(ROOT, "julia", "AlphaGoB200.jl")).read()
    for jname, cname in (("AgzConfig", "agz_config"), ("AgzGameHeader", "agz_game_header"), ("AgzProgress", "agz_progress")):
        m = re.search(r"^struct %s\n(.*?)^end" % jname, jl, flags=re.S | re.M)
        assert m, jname
        got = re.findall(r"(\w+)::(\w+)", m.group(1))
        want = [(f, JULIA_TYPES[t]) for f, t, arr in fields[cname]]
        assert got == want, "%s differs from %s" % (jname, cname)


def test_every_exported_symbol_is_declared_in_the_header_and_bound():
    src = open(os.path.join(ROOT, "include", "agz.h")).read()
    declared = set(re.findall(r"\b(agz_\w+)\s*\(", re.sub(r"/\*.*?\*/", "", src, flags=re.S)))
    assert declared == set(B.SYMBOLS), (declared - set(B.SYMBOLS), set(B.SYMBOLS) - declared)
    jl = open(os.path.join(ROOT, "julia", "AlphaGoB200.jl")).read()
    used = set(re.findall(r"\(:(agz_\w+), libagz\)", jl))
    assert used <= declared, used - declared
