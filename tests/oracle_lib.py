"""ctypes wrapper of the CPU oracle (oracle/liboracle.so). TEST INFRASTRUCTURE: imported by tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")
L_ORDER = 2**252 + 27742317777372353535851937790883648493
R_MONT = (1 << 256) % L_ORDER
R_INV = pow(R_MONT, -1, L_ORDER)
COO_DTYPE = np.dtype([("row", "<u8"), ("col", "<u8"), ("val", "u1", (32,))])

_lib = None


def build():
    subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        for f in ("orc_build_point_mult", "orc_build_point_add", "orc_build_custom", "orc_run_flow", "orc_transcript_new"):
            getattr(_lib, f).restype = C.c_void_p
        _lib.orc_flow_get.restype = C.c_uint64
        _lib.orc_timer_name.restype = C.c_char_p
    return _lib


def le32(x):
    return int(x % L_ORDER).to_bytes(32, "little")


def ints_to_bytes(xs):
    return b"".join(le32(x) for x in xs)


def bytes_to_ints(b):
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def to_mont_bytes(xs):
    return b"".join(((x * R_MONT) % L_ORDER).to_bytes(32, "little") for x in xs)


def from_mont_bytes(b):
    return [(int.from_bytes(b[i:i + 32], "little") * R_INV) % L_ORDER for i in range(0, len(b), 32)]


class Built:
    """BuiltInstance handle (instance triples + the three assignments)."""

    def __init__(self, h):
        assert h, "oracle builder failed"
        self.h = C.c_void_p(h)
        info = (C.c_uint64 * 7)()
        lib().orc_built_info(self.h, info)
        self.num_cons, self.num_vars, self.num_inputs, self.nnz_param, self.nA, self.nB, self.nC = list(info)
        self.dims = (self.num_cons, self.num_vars, self.num_inputs, self.nnz_param)

    def arrays(self):
        A = np.zeros(self.nA, COO_DTYPE)
        B = np.zeros(self.nB, COO_DTYPE)
        Cm = np.zeros(self.nC, COO_DTYPE)
        bufs = [C.create_string_buffer(32 * self.num_vars) for _ in range(3)]
        inputs = C.create_string_buffer(max(32 * self.num_inputs, 1))
        lib().orc_built_copy(self.h, A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p), Cm.ctypes.data_as(C.c_void_p), bufs[0],
                             bufs[1], bufs[2], inputs)
        return A, B, Cm, bufs[0].raw, bufs[1].raw, bufs[2].raw, inputs.raw[: 32 * self.num_inputs]

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_built_free(self.h)
            self.h = None


def build_point_mult(weights, px, py):
    m = len(weights)
    w = (C.c_uint64 * (2 * m))(*[x for ww in weights for x in (ww & (2**64 - 1), ww >> 64)])
    return Built(lib().orc_build_point_mult(C.c_uint64(m), w, px, py))


def build_point_add(px, py, rx, ry, rz):
    n = len(rz)
    return Built(lib().orc_build_point_add(C.c_uint64(n), px, py, rx, ry, (C.c_int64 * n)(*rz)))


def build_custom(num_cons, num_vars, num_inputs, nnz_param, A, B, Cm, vars_para, vars_input, vars_, inputs):
    A, B, Cm = (np.ascontiguousarray(x, dtype=COO_DTYPE) for x in (A, B, Cm))
    return Built(lib().orc_build_custom(C.c_uint64(num_cons), C.c_uint64(num_vars), C.c_uint64(num_inputs), C.c_uint64(nnz_param),
                                        A.ctypes.data_as(C.c_void_p), C.c_uint64(len(A)), B.ctypes.data_as(C.c_void_p), C.c_uint64(len(B)),
                                        Cm.ctypes.data_as(C.c_void_p), C.c_uint64(len(Cm)), vars_para, vars_input, vars_, inputs))


class Flow:
    def __init__(self, built, seed_q, seed_p, verify=True, threads=None):
        threads = threads or os.cpu_count() or 1
        h = lib().orc_run_flow(built.h, seed_q, seed_p, C.c_int(1 if verify else 0), C.c_int(threads))
        assert h, "oracle flow failed"
        self.h = C.c_void_p(h)
        self.verified = bool(lib().orc_flow_verified(self.h))
        self.proof, self.comm, self.comm_vars_para, self.comm_vars_input, self.comm_vars = (self._get(i) for i in range(5))
        n = lib().orc_num_timers()
        tm = (C.c_double * (n + 3))()
        lib().orc_flow_times(self.h, tm)
        self.times = {lib().orc_timer_name(i).decode(): tm[i] for i in range(n + 3)}
        lib().orc_flow_free(self.h)
        self.h = None

    def _get(self, which):
        n = lib().orc_flow_get(self.h, which, None, C.c_uint64(0))
        buf = C.create_string_buffer(max(n, 1))
        lib().orc_flow_get(self.h, which, buf, C.c_uint64(n))
        return buf.raw[:n]


def verify(dims, proof, comm, inputs, com_1, com_2):
    num_cons, num_vars, num_inputs, nnz = dims
    return lib().orc_verify(C.c_uint64(num_cons), C.c_uint64(num_vars), C.c_uint64(num_inputs), C.c_uint64(nnz), proof,
                            C.c_uint64(len(proof)), comm, C.c_uint64(len(comm)), inputs, com_1, com_2, C.c_uint64(len(com_1) // 32))


# ---- small helpers over Montgomery arrays (kernel-level parity) ----
def eq_evals(r_ints):
    out = C.create_string_buffer(32 << len(r_ints))
    lib().orc_eq_evals_mont(to_mont_bytes(r_ints), C.c_uint64(len(r_ints)), out)
    return from_mont_bytes(out.raw)


def cubic_round(A, B, Cc, D):
    out = C.create_string_buffer(96)
    lib().orc_cubic_round_mont(to_mont_bytes(A), to_mont_bytes(B), to_mont_bytes(Cc), to_mont_bytes(D), C.c_uint64(len(A)), out)
    return from_mont_bytes(out.raw)


def quad_round(A, B):
    out = C.create_string_buffer(64)
    lib().orc_quad_round_mont(to_mont_bytes(A), to_mont_bytes(B), C.c_uint64(len(A)), out)
    return from_mont_bytes(out.raw)


def cubic3_round(A, B, Cc):
    out = C.create_string_buffer(96)
    lib().orc_cubic3_round_mont(to_mont_bytes(A), to_mont_bytes(B), to_mont_bytes(Cc), C.c_uint64(len(A)), out)
    return from_mont_bytes(out.raw)


def bind_top(Z, r):
    buf = C.create_string_buffer(to_mont_bytes(Z), 32 * len(Z))
    lib().orc_bind_top_mont(buf, C.c_uint64(len(Z)), to_mont_bytes([r]))
    return from_mont_bytes(buf.raw[: 16 * len(Z)])


def spmv(M, num_rows, num_cols, z, transposed=False):
    M = np.ascontiguousarray(M, dtype=COO_DTYPE)
    out = C.create_string_buffer(32 * (num_cols if transposed else num_rows))
    fn = lib().orc_spmv_t_mont if transposed else lib().orc_spmv_mont
    ok = fn(M.ctypes.data_as(C.c_void_p), C.c_uint64(len(M)), C.c_uint64(num_rows), C.c_uint64(num_cols), to_mont_bytes(z), out)
    assert ok == 1
    return from_mont_bytes(out.raw)


def hyrax_commit(Z, label, blinds=None, threads=8):
    n = len(Z)
    ell = n.bit_length() - 1
    Lr = 1 << (ell // 2)
    out = C.create_string_buffer(32 * Lr)
    lib().orc_hyrax_commit_mont(to_mont_bytes(Z), C.c_uint64(n), to_mont_bytes(blinds) if blinds is not None else None, label,
                                C.c_int(threads), out)
    return out.raw


def bound(Z, Lvec):
    n = len(Z)
    ell = n.bit_length() - 1
    R = 1 << (ell - ell // 2)
    out = C.create_string_buffer(32 * R)
    lib().orc_bound_mont(to_mont_bytes(Z), C.c_uint64(n), to_mont_bytes(Lvec), out)
    return from_mont_bytes(out.raw)


def derive_gens(label, n):
    out = C.create_string_buffer(32 * (n + 1))
    lib().orc_derive_gens(label, C.c_uint64(n), out)
    return out.raw


def msm(scalars_ints, points_bytes):
    out = C.create_string_buffer(32)
    ok = lib().orc_msm(C.c_uint64(len(scalars_ints)), ints_to_bytes(scalars_ints), points_bytes, out)
    assert ok == 1
    return out.raw
