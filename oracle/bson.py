"""Minimal BSON reader for the reference's models/**/agz_*.bson (BSON.jl array encoding).
TEST INFRASTRUCTURE.  Only what is needed to pull Float32/Float64 arrays out."""
import struct

import numpy as np


def _cstring(d, o):
    e = d.index(b"\x00", o)
    return d[o:e].decode("utf8", "replace"), e + 1


def _parse_doc(d, o, as_list=False):
    size = struct.unpack_from("<i", d, o)[0]
    end = o + size - 1
    o += 4
    out = [] if as_list else {}
    while o < end:
        t = d[o]
        o += 1
        name, o = _cstring(d, o)
        if t == 0x01:
            v = struct.unpack_from("<d", d, o)[0]; o += 8
        elif t == 0x02:
            n = struct.unpack_from("<i", d, o)[0]; o += 4
            v = d[o:o + n - 1].decode("utf8", "replace"); o += n
        elif t == 0x03:
            v, o = _parse_doc(d, o)
        elif t == 0x04:
            v, o = _parse_doc(d, o, as_list=True)
        elif t == 0x05:
            n = struct.unpack_from("<i", d, o)[0]; o += 5
            v = d[o:o + n]; o += n
        elif t == 0x08:
            v = bool(d[o]); o += 1
        elif t == 0x0A:
            v = None
        elif t == 0x10:
            v = struct.unpack_from("<i", d, o)[0]; o += 4
        elif t == 0x12:
            v = struct.unpack_from("<q", d, o)[0]; o += 8
        else:
            raise ValueError("unsupported BSON type 0x%02x at %d" % (t, o))
        if as_list:
            out.append(v)
        else:
            out[name] = v
    return out, end + 1


def _decode(v):
    """Turn BSON.jl tagged documents into numpy arrays / python containers."""
    if isinstance(v, dict):
        if v.get("tag") == "array" and isinstance(v.get("type"), dict):
            tname = v["type"].get("name", [])
            dt = {"Float32": "<f4", "Float64": "<f8", "Int64": "<i8", "UInt8": "u1", "Bool": "u1"}.get(tname[-1] if tname else "")
            if dt is not None and isinstance(v.get("data"), (bytes, bytearray)):
                shape = tuple(int(s) for s in v["size"])
                return np.frombuffer(v["data"], dtype=dt).reshape(shape, order="F").copy()
        return {k: _decode(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_decode(x) for x in v]
    return v


def load(path):
    d = open(path, "rb").read()
    doc, _ = _parse_doc(d, 0)
    return _decode(doc)


def find_arrays(v, out=None):
    """Depth-first list of every numpy array in a decoded document, in file order."""
    if out is None:
        out = []
    if isinstance(v, np.ndarray):
        out.append(v)
    elif isinstance(v, dict):
        if v.get("tag") == "array" and "data" in v and isinstance(v["data"], list):
            for x in v["data"]:
                find_arrays(x, out)
        else:
            for x in v.values():
                find_arrays(x, out)
    elif isinstance(v, list):
        for x in v:
            find_arrays(x, out)
    return out
