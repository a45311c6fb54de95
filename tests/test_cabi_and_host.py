"""CPU: the C-ABI shared library loads and exports every symbol include/vpin_b200.h declares (no compute calls without a
GPU); the product fails loudly without a CUDA device; host-side logic (witness file format, shape tables)."""
import ctypes as C
import json
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "vpin_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vpin_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from vpin_b200 import api
    lib = api.lib()
    names = declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/vpin_b200.h but not exported: {missing}"


def test_dims_entry_points_need_no_gpu():
    """vpin_point_mult_dims / vpin_point_add_dims mirror the hard-coded nnz tables (point_mult.rs:27-67, point_addition.rs:38-70)"""
    from vpin_b200 import api
    lib = api.lib()
    d = (C.c_uint64 * 4)()
    lib.vpin_point_mult_dims(C.c_uint64(18), d)
    assert list(d)[:3] == [18 * 3464, 18 * 3466 + 1, 1]
    lib.vpin_point_mult_dims(C.c_uint64(178), d)
    assert list(d) == [616592, 616949, 1, 737600]
    lib.vpin_point_add_dims(C.c_uint64(2144), d)
    assert list(d) == [21440, 32161, 0, 64320]
    # every named shape: the nnz parameter must round to the same power of two as the true padded max nnz (41n+12 per mult)
    from vpin_b200 import workloads as W
    for tag, (m, n_add) in W.SHAPES.items():
        if m == 0:  # LeNet's pooling layers have point additions only
            continue
        lib.vpin_point_mult_dims(C.c_uint64(m), d)
        true_nnz = m * (41 * 128 + 12)
        assert (int(d[3]) - 1).bit_length() == (true_nnz - 1).bit_length(), tag


def test_no_cpu_fallback():
    """without a CUDA device the context cannot be created: the product path never computes on the CPU"""
    import torch
    from vpin_b200 import api
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.VpinError) as e:
        api.Context(0)
    assert e.value.name == "CudaError"


def test_tape_init_rejects_non_canonical_seed():
    from vpin_b200 import api
    with pytest.raises(api.VpinError):
        api.RandomTape(b"proof", bytes([0xFF] * 32))
    api.RandomTape(b"proof", bytes(32))


def test_product_sources_do_not_touch_the_oracle():
    """the oracle is test infrastructure: nothing under vpin_b200/ may import, include or link it"""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "vpin_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"oracle_lib|liboracle|\boracle/|import pyref|from oracle", txt):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_witness_json_round_trip(tmp_path):
    """rust_files/<tag>/{pointMult,pointAdd}/*.json keep the reference's format (load_data.rs:5-62, load_data_add.rs:5-102)"""
    from vpin_b200 import workloads as W
    mult = W.synth_point_mult(5)
    add = W.synth_point_add(6, infinity_every=4)
    W.write_rust_files(str(tmp_path), "conv3", mult=mult, add=add)
    d = tmp_path / "rust_files" / "conv3"
    weights = json.load(open(d / "pointMult" / "weight.json"))
    assert all(isinstance(w, str) and w.isdigit() for w in weights)  # decimal strings parsed as u128
    px = json.load(open(d / "pointMult" / "point_mult_px_byte.json"))
    assert len(px) == 5 and all(len(row) == 32 and all(0 <= b < 256 for b in row) for row in px)
    rz = json.load(open(d / "pointAdd" / "point_add_rz_byte.json"))
    assert rz == [1, 0, 0, 0, 1, 0]
    assert W.load_point_mult(str(tmp_path), "conv3") == mult
    got = W.load_point_add(str(tmp_path), "conv3")
    assert got[:4] == add[:4] and got[4] == add[4]
    # points at infinity are written as (0, 0)
    assert add[2][:32] == bytes(32) and add[3][:32] == bytes(32)


def test_synthetic_points_are_on_vpins_curve():
    from vpin_b200 import workloads as W
    _, px, py = W.synth_point_mult(9)
    for i in range(9):
        x = int.from_bytes(px[32 * i:32 * i + 32], "little")
        y = int.from_bytes(py[32 * i:32 * i + 32], "little")
        assert (y * y - (x * x * x + W.CURVE_A * x + W.CURVE_B)) % W.FIELD == 0
    assert W.ec_mul(W.ORDER, W.GEN) is None
    assert W.FIELD == 2**252 + 27742317777372353535851937790883648493  # the curve's base field is Spartan's scalar field


def test_native_witness_loader_matches_the_python_one(tmp_path):
    """(f4) vpin_load_point_mult / vpin_load_point_add (C++, VP/load_data.rs:5-62, load_data_add.rs:5-102) return the same bytes
    as the Python loader from the reference's JSON files, and again from the witness.bin sidecar; malformed input is an error
    code, not an abort."""
    import pytest
    from vpin_b200 import api, workloads as W
    mult = W.synth_point_mult(37)
    mult = ([0, 1, (1 << 128) - 1] + list(mult[0][3:]), mult[1], mult[2])   # u128 edge values
    add = W.synth_point_add(41, infinity_every=5)
    root = str(tmp_path)
    W.write_rust_files(root, "netX", mult=mult, add=add)
    assert W.load_point_mult_native(root, "netX") == W.load_point_mult(root, "netX") == mult
    got = W.load_point_add_native(root, "netX")
    assert got[:4] == add[:4] and got[4] == add[4]
    # sidecar: written once, preferred afterwards (remove the JSON to prove it is what gets read)
    assert W.witness_json_to_bin(root, "netX") == 3
    for sub, names in (("pointMult", ["weight.json", "point_mult_px_byte.json", "point_mult_py_byte.json"]),
                       ("pointAdd", ["point_add_px_byte.json", "point_add_py_byte.json", "point_add_rx_byte.json", "point_add_ry_byte.json",
                                     "point_add_rz_byte.json"])):
        for n in names:
            os.remove(os.path.join(root, "rust_files", "netX", sub, n))
    assert W.load_point_mult_native(root, "netX") == mult
    got = W.load_point_add_native(root, "netX")
    assert got[:4] == add[:4] and got[4] == add[4]
    # tolerated like serde_json + as_i64: whitespace, short rows (zero padded), non-integer entries skipped
    d = os.path.join(root, "rust_files", "netY", "pointMult")
    os.makedirs(d)
    open(os.path.join(d, "weight.json"), "w").write(' [ "7" ,\n "340282366920938463463374607431768211455" ] ')
    open(os.path.join(d, "point_mult_px_byte.json"), "w").write("[[1, 2, 3], [255, 1.5, 4]]")
    open(os.path.join(d, "point_mult_py_byte.json"), "w").write("[[], [9]]")
    w, px, py = W.load_point_mult_native(root, "netY")
    assert w == [7, (1 << 128) - 1]
    assert px == bytes([1, 2, 3]) + bytes(29) + bytes([255, 4]) + bytes(30) and py == bytes(32) + bytes([9]) + bytes(31)
    # errors
    with pytest.raises(api.VpinError) as e:
        W.load_point_mult_native(root, "missing")
    assert e.value.name == "IoError"
    open(os.path.join(d, "weight.json"), "w").write('["340282366920938463463374607431768211456", "1"]')   # 2^128
    with pytest.raises(api.VpinError) as e:
        W.load_point_mult_native(root, "netY")
    assert e.value.name == "InvalidScalar"
    open(os.path.join(d, "weight.json"), "w").write('["7", "8", "9"]')
    with pytest.raises(api.VpinError) as e:
        W.load_point_mult_native(root, "netY")
    assert e.value.name == "SizeMismatch"
    open(os.path.join(d, "weight.json"), "w").write('["7", "8"')
    with pytest.raises(api.VpinError) as e:
        W.load_point_mult_native(root, "netY")
    assert e.value.name == "IoError"
