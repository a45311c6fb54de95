"""Synthetic vPIN witnesses of the named shapes.

The reference's witness VALUES are random (client keys, os.urandom: src/convolution/Client.py:23,
Server.py:264), so only the shapes are named.  This module reproduces what the Python side writes into
rust_files/<tag>/{pointMult,pointAdd}/*.json (src/convolution/Server.py:311-417) — 32-byte little-endian
affine coordinates on vPIN's ElGamal curve and decimal-string u128 weights — from a fixed seed, and can
also read/write those JSON files unchanged (vPIN_proof_generation/src/load_data.rs, load_data_add.rs).
"""
import json
import os

import numpy as np

# src/convolution/Client.py:138-143 (curveE2Info): y^2 = x^3 + a x + b over F_l, l = ristretto255 group order
FIELD = 7237005577332262213973186563042994240857116359379907606001950938285454250989
CURVE_A = 3491403595575449084947959021303599933011749826127899762162894550148391771037
CURVE_B = 3633908682298454119909199192149978293706667958442512986315258451820769071958
GEN = (4561981307020378385254256586024830594940985765081274686120783167106442831732,
       684120277165286233470758410892647831027470652988879249692043589061244861334)
ORDER = 7237005577332262213973186563042994240704759454384003648147593987722918659549

SEED = 0x7650494E

# (point multiplications, point additions) per named config — SURVEY.md section 8d table
SHAPES = {
    "conv3": (18, 16), "conv5": (50, 48), "conv7": (98, 96),
    "A": (178, 2144), "B": (210, 2208), "C": (562, 2144), "D": (594, 2208), "E": (658, 2336),
    # LeNet, per layer (point multiplications, point additions): the slices of src/LeNet/Server.py:690-698 and :753-761
    "L1": (300, 288), "L2": (0, 7056), "L3": (800, 768), "L4": (0, 2400), "L5": (6000, 5760), "L6": (240, 406), "L7": (168, 186),
}
LENET_LAYERS = ("L1", "L2", "L3", "L4", "L5", "L6", "L7")


def ec_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    (x1, y1), (x2, y2) = p, q
    if x1 == x2:
        if (y1 + y2) % FIELD == 0:
            return None
        lam = (3 * x1 * x1 + CURVE_A) * pow(2 * y1, -1, FIELD) % FIELD
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, FIELD) % FIELD
    x3 = (lam * lam - x1 - x2) % FIELD
    return x3, (lam * (x1 - x3) - y1) % FIELD


def ec_mul(k, p):
    acc = None
    while k:
        if k & 1:
            acc = ec_add(acc, p)
        p = ec_add(p, p)
        k >>= 1
    return acc


def _rand_int(rng, bits):
    nbytes = (bits + 7) // 8
    return int.from_bytes(rng.bytes(nbytes), "little") >> (nbytes * 8 - bits)


def _rand_points(rng, n):
    """n distinct-looking points k*G; walks by a random stride so generation stays O(n) additions."""
    base = ec_mul(1 + _rand_int(rng, 250) % (ORDER - 2), GEN)
    stride = ec_mul(1 + _rand_int(rng, 250) % (ORDER - 2), GEN)
    out = []
    cur = base
    for _ in range(n):
        out.append(cur)
        cur = ec_add(cur, stride)
    return out


def le32(x):
    return int(x).to_bytes(32, "little")


def synth_point_mult(m, seed=SEED, weight_bits=112, small_weights=False):
    """Returns (weights: list[int] < 2^128, px: bytes m*32, py: bytes m*32).
    small_weights=True mimics conv filters (values 0/1/2, src/convolution/Server.py:452-469);
    otherwise 112-bit weights as produced by the FC layers' pf() (src/cnn_networks/Server.py:406-411)."""
    rng = np.random.default_rng(seed)
    pts = _rand_points(rng, m)
    if small_weights:
        weights = [int(v) for v in rng.integers(0, 3, size=m)]
    else:
        weights = [_rand_int(rng, weight_bits) for _ in range(m)]
    px = b"".join(le32(p[0]) for p in pts)
    py = b"".join(le32(p[1]) for p in pts)
    return weights, px, py


def synth_point_add(n, seed=SEED + 1, infinity_every=0):
    """Returns (px, py, rx, ry: bytes n*32 each, rz: list[int]).  rz=1 marks R = infinity, for which the
    Python side writes rx = ry = 0 (src/convolution/Server.py:373-381)."""
    rng = np.random.default_rng(seed)
    P = _rand_points(rng, n)
    R = _rand_points(rng, n)
    rz = [1 if (infinity_every and i % infinity_every == 0) else 0 for i in range(n)]
    px = b"".join(le32(p[0]) for p in P)
    py = b"".join(le32(p[1]) for p in P)
    rx = b"".join(le32(0 if z else r[0]) for r, z in zip(R, rz))
    ry = b"".join(le32(0 if z else r[1]) for r, z in zip(R, rz))
    return px, py, rx, ry, rz


def write_rust_files(root, tag, mult=None, add=None):
    """Writes the reference's JSON witness files (format unchanged)."""
    if mult is not None:
        weights, px, py = mult
        d = os.path.join(root, "rust_files", tag, "pointMult")
        os.makedirs(d, exist_ok=True)
        json.dump([str(w) for w in weights], open(os.path.join(d, "weight.json"), "w"))
        json.dump([list(px[32 * i:32 * i + 32]) for i in range(len(weights))], open(os.path.join(d, "point_mult_px_byte.json"), "w"))
        json.dump([list(py[32 * i:32 * i + 32]) for i in range(len(weights))], open(os.path.join(d, "point_mult_py_byte.json"), "w"))
    if add is not None:
        px, py, rx, ry, rz = add
        d = os.path.join(root, "rust_files", tag, "pointAdd")
        os.makedirs(d, exist_ok=True)
        n = len(rz)
        for name, buf in (("px", px), ("py", py), ("rx", rx), ("ry", ry)):
            json.dump([list(buf[32 * i:32 * i + 32]) for i in range(n)], open(os.path.join(d, f"point_add_{name}_byte.json"), "w"))
        json.dump([int(z) for z in rz], open(os.path.join(d, "point_add_rz_byte.json"), "w"))


def _rows_to_bytes(rows):
    out = bytearray()
    for row in rows:
        b = bytearray(32)
        for j, v in enumerate(row):
            b[j] = int(v) & 0xFF
        out += b
    return bytes(out)


def load_point_mult(root, tag):
    """load_data.rs:5-62"""
    d = os.path.join(root, "rust_files", tag, "pointMult")
    weights = [int(s) for s in json.load(open(os.path.join(d, "weight.json")))]
    px = _rows_to_bytes(json.load(open(os.path.join(d, "point_mult_px_byte.json"))))
    py = _rows_to_bytes(json.load(open(os.path.join(d, "point_mult_py_byte.json"))))
    return weights, px, py


def load_point_add(root, tag):
    """load_data_add.rs:5-102"""
    d = os.path.join(root, "rust_files", tag, "pointAdd")
    bufs = [_rows_to_bytes(json.load(open(os.path.join(d, f"point_add_{n}_byte.json")))) for n in ("px", "py", "rx", "ry")]
    rz = [int(v) for v in json.load(open(os.path.join(d, "point_add_rz_byte.json")))]
    return (*bufs, rz)


def load_point_mult_native(root, tag):
    """the same through the library's native loader (vpin_load_point_mult: JSON or the witness.bin sidecar)"""
    import ctypes as C
    from vpin_b200 import api
    L = api.lib()
    n = C.c_uint64()
    st = L.vpin_load_point_mult(root.encode(), tag.encode(), C.c_uint64(0), C.byref(n), None, None, None)
    if st != 0:
        raise api.VpinError(st, L.vpin_witness_last_error().decode())
    m = n.value
    w = (C.c_uint64 * (2 * m))()
    px, py = C.create_string_buffer(32 * m), C.create_string_buffer(32 * m)
    st = L.vpin_load_point_mult(root.encode(), tag.encode(), C.c_uint64(m), C.byref(n), w, px, py)
    if st != 0:
        raise api.VpinError(st, L.vpin_witness_last_error().decode())
    return [int(w[2 * i]) | (int(w[2 * i + 1]) << 64) for i in range(m)], px.raw[: 32 * m], py.raw[: 32 * m]


def load_point_add_native(root, tag):
    import ctypes as C
    from vpin_b200 import api
    L = api.lib()
    n = C.c_uint64()
    st = L.vpin_load_point_add(root.encode(), tag.encode(), C.c_uint64(0), C.byref(n), None, None, None, None, None)
    if st != 0:
        raise api.VpinError(st, L.vpin_witness_last_error().decode())
    m = n.value
    bufs = [C.create_string_buffer(32 * m) for _ in range(4)]
    rz = (C.c_int64 * m)()
    st = L.vpin_load_point_add(root.encode(), tag.encode(), C.c_uint64(m), C.byref(n), *bufs, rz)
    if st != 0:
        raise api.VpinError(st, L.vpin_witness_last_error().decode())
    return (*[b.raw[: 32 * m] for b in bufs], [int(z) for z in rz])


def witness_json_to_bin(root, tag):
    """writes the witness.bin sidecars (vpin_witness_json_to_bin); returns the bit mask of what was written"""
    import ctypes as C
    from vpin_b200 import api
    L = api.lib()
    out = C.c_uint32()
    st = L.vpin_witness_json_to_bin(root.encode(), tag.encode(), C.byref(out))
    if st != 0:
        raise api.VpinError(st, L.vpin_witness_last_error().decode())
    return out.value


def tape_seeds():
    """TEST-ONLY deterministic stand-ins for the two OsRng scalars: 64 B of SHAKE256 -> mod l.
    They make proofs byte-comparable with the oracle; a real prover passes NULL seeds (the library then draws them from the
    OS) - reusing a seed for a different witness reuses blinds and nonces and leaks the witness."""
    import hashlib
    L = 2**252 + 27742317777372353535851937790883648493
    q = int.from_bytes(hashlib.shake_256(b"vpin-b200/tape/0x02").digest(64), "little") % L
    p = int.from_bytes(hashlib.shake_256(b"vpin-b200/tape/proof").digest(64), "little") % L
    return le32(q), le32(p)
