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


# ------------------------------------------------------------------ writing (save_model, src/train.jl:14-35)
def _cstr_b(name):
    return name.encode("utf8") + b"\x00"


def _enc(value):
    """BSON element (type byte, payload) for the few value kinds BSON.jl's array documents use."""
    if isinstance(value, dict):
        return 0x03, _enc_doc(value)
    if isinstance(value, list):
        return 0x04, _enc_doc({str(i): v for i, v in enumerate(value)})
    if isinstance(value, (bytes, bytearray)):
        return 0x05, struct.pack("<i", len(value)) + b"\x00" + bytes(value)
    if isinstance(value, str):
        b = value.encode("utf8") + b"\x00"
        return 0x02, struct.pack("<i", len(b)) + b
    if isinstance(value, (int, np.integer)):
        return 0x12, struct.pack("<q", int(value))
    raise TypeError("unsupported value %r" % type(value))


def _enc_doc(d):
    body = b""
    for k, v in d.items():
        t, payload = _enc(v)
        body += bytes([t]) + _cstr_b(k) + payload
    return struct.pack("<i", len(body) + 5) + body + b"\x00"


def _array_doc(a):
    a = np.asarray(a, np.float32)
    return {"tag": "array", "type": {"tag": "datatype", "params": [], "name": ["Core", "Float32"]},
            "size": [int(x) for x in a.shape], "data": a.flatten(order="F").tobytes()}


def save_weight_list(path, name, arrays):
    """One `@save path name` of save_model: a top-level document {name: [array documents]} in BSON.jl's array encoding."""
    with open(path, "wb") as f:
        f.write(_enc_doc({name: [_array_doc(a) for a in arrays]}))


def save_model(nn, models_dir, engine=None):
    """save_model(nn) (src/train.jl:14-35): weights/agz_{base,value,policy}.bson hold the three Flux `params` lists in the
    reference's layout (load_reference_model reads them back).  The reference also dumps the whole Flux structs (agz_*.bson); the
    files of that name written HERE are an engine-private format carrying what this engine needs from them -- the BatchNorm running
    statistics per chain as (mean, VARIANCE): statistics held as a moving standard deviation (bn_mode BN_STD, the shipped
    models/agz_*.bson) are squared on the way out (eps of that convention is 1e-8 vs 1e-5: a 1e-5-relative difference in the folded
    scale).  With `engine`, the current (trained) parameters are read back from it first."""
    import os
    if engine is not None:
        pull_from_engine(nn, engine)
    os.makedirs(os.path.join(models_dir, "weights"), exist_ok=True)
    for k, (chain, var) in enumerate((("base", "bn_weights"), ("value", "val_weights"), ("policy", "pol_weights"))):
        save_weight_list(os.path.join(models_dir, "weights", "agz_%s.bson" % chain), var, nn.params[k])
        var = np.asarray(nn.bn_sigma[k], np.float32) ** 2 if nn.bn_mode == B.BN_STD else nn.bn_sigma[k]
        save_weight_list(os.path.join(models_dir, "agz_%s.bson" % chain), chain + "_bn_stats", [nn.bn_mu[k], var])


def load_saved_model(models_dir, nn):
    """Inverse of save_model above."""
    lists, mus, sigmas = [], [], []
    for chain in ("base", "value", "policy"):
        lists.append(bson_arrays(os_join(models_dir, "weights", "agz_%s.bson" % chain)))
        st = bson_arrays(os_join(models_dir, "agz_%s.bson" % chain))
        mus.append(st[0].ravel())
        sigmas.append(st[1].ravel())
    nn.load_params(base=lists[0], value=lists[1], policy=lists[2], bn_mu=mus, bn_sigma=sigmas, bn_mode=B.BN_VAR_EPS)
    return nn


def os_join(*parts):
    import os
    return os.path.join(*parts)


def pull_from_engine(nn, engine):
    """Copy the engine's current parameters / running statistics (after agz_train_step) into an api.NeuralNet."""
    for k in range(3):
        flat = engine.net_get_params(k)
        out, o = [], 0
        for a in nn.params[k]:
            n = int(np.prod(a.shape))
            out.append(flat[o:o + n].reshape(a.shape, order="F").copy())
            o += n
        nn.params[k] = out
        mu, sg, mode = engine.net_get_bn_stats(k)
        nn.bn_mu[k], nn.bn_sigma[k] = mu, sg
        nn.bn_mode = mode
    nn._version += 1
    engine._nn_token = (id(nn), nn._version)
    return nn
