"""Reading the reference's saved networks: `save_model` (src/train.jl:14-35) writes
models/weights/agz_{base,value,policy}.bson (the three Flux `params` lists) and models/agz_*.bson (whole structs, which
also carry the BatchNorm running statistics).  This is a minimal BSON.jl array reader -- no compute, host side only."""
import struct

import numpy as np

from . import binding as B

_DTYPES = {"Float32": "<f4", "Float64": "<f8", "Int64": "<i8", "Int32": "<i4", "UInt8": "u1", "Bool": "u1"}


def _cstr(d, o):
    e = d.index(b"\x00", o)
    return d[o:e].decode("utf8", "replace"), e + 1


def _doc(d, o, as_list=False):
    size = struct.unpack_from("<i", d, o)[0]
    end = o + size - 1
    o += 4
    out = [] if as_list else {}
    while o < end:
        t = d[o]
        name, o = _cstr(d, o + 1)
        if t == 0x01:
            v, o = struct.unpack_from("<d", d, o)[0], o + 8
        elif t == 0x02:
            n = struct.unpack_from("<i", d, o)[0]
            v, o = d[o + 4:o + 4 + n - 1].decode("utf8", "replace"), o + 4 + n
        elif t in (0x03, 0x04):
            v, o = _doc(d, o, as_list=(t == 0x04))
        elif t == 0x05:
            n = struct.unpack_from("<i", d, o)[0]
            v, o = d[o + 5:o + 5 + n], o + 5 + n
        elif t == 0x08:
            v, o = bool(d[o]), o + 1
        elif t == 0x0A:
            v = None
        elif t == 0x10:
            v, o = struct.unpack_from("<i", d, o)[0], o + 4
        elif t == 0x12:
            v, o = struct.unpack_from("<q", d, o)[0], o + 8
        else:
            raise ValueError("unsupported BSON element type 0x%02x" % t)
        if as_list:
            out.append(v)
        else:
            out[name] = v
    return out, end + 1


def _arrays(v, out):
    """Every dense numeric array of a BSON.jl document, in file order (column-major -> numpy order='F')."""
    if isinstance(v, dict):
        if v.get("tag") == "array" and isinstance(v.get("type"), dict) and isinstance(v.get("data"), (bytes, bytearray)):
            name = v["type"].get("name") or [""]
            dt = _DTYPES.get(name[-1])
            if dt is not None:
                out.append(np.frombuffer(v["data"], dtype=dt).reshape(tuple(int(s) for s in v["size"]), order="F").copy())
                return
        for x in v.values():
            _arrays(x, out)
    elif isinstance(v, list):
        for x in v:
            _arrays(x, out)


def bson_arrays(path):
    doc, _ = _doc(open(path, "rb").read(), 0)
    out = []
    _arrays(doc, out)
    return out


def load_reference_model(models_dir, nn):
    """Fill an api.NeuralNet (tower_height 0 for the shipped 9x9 `agz` net) from a directory laid out like the reference's
    models/: weights/agz_{base,value,policy}.bson + agz_{base,value,policy}.bson (BatchNorm mu / moving std)."""
    lists, mus, sigmas = [], [], []
    for name in ("base", "value", "policy"):
        lists.append(bson_arrays("%s/weights/agz_%s.bson" % (models_dir, name)))
        full = bson_arrays("%s/agz_%s.bson" % (models_dir, name))
        nbn = (len(lists[-1]) - (0 if name == "base" else (4 if name == "value" else 2))) // 4 if name == "base" else 1
        # the struct files list, per BatchNorm layer, mu then sigma (untracked arrays) before the tracked parameters
        mus.append(np.concatenate([full[2 * k].ravel() for k in range(nbn)]))
        sigmas.append(np.concatenate([full[2 * k + 1].ravel() for k in range(nbn)]))
    nn.load_params(base=lists[0], value=lists[1], policy=lists[2], bn_mu=mus, bn_sigma=sigmas, bn_mode=B.BN_STD)
    return nn
