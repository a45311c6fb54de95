"""ctypes binding of libvpin_b200.so (include/vpin_b200.h) mirroring the reference's Rust interface for the prover
path: SNARKGens::new, Instance::new, SNARK::encode, DensePolynomial::commit, my_dense_mlpoly_commit, my_lib_prove
(Spartan/src/lib.rs:138-358, vPIN_proof_generation/src/commit_test.rs:27-133).

This is the host side a Rust shim would replace one-for-one. There is NO CPU fallback: the
library must be built (`python -c "import __graft_entry__ as g; g.build()"`) and a CUDA device must be present, or the
calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VPIN_B200_LIB", os.path.join(_HERE, "libvpin_b200.so"))

COO_DTYPE = np.dtype([("row", "<u8"), ("col", "<u8"), ("val", "u1", (32,))])

STATUS = {
    0: "OK", 1: "InvalidScalar", 2: "InvalidIndex", 3: "InvalidNumberOfInputs", 4: "SizeMismatch", 5: "BufferTooSmall",
    6: "CudaError", 7: "OutOfMemory", 8: "ProverAssertion", 9: "BadArgument", 10: "IoError",
}


class VpinError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code
        self.name = STATUS.get(code, str(code))


_lib = None


def lib():
    """Loads the CUDA library; raises (loudly) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with __graft_entry__.build(); there is no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        L = _lib
        vp, u64 = C.c_void_p, C.c_uint64
        L.vpin_witness_last_error.restype = C.c_char_p
        L.vpin_last_error.restype = C.c_char_p
        L.vpin_last_error.argtypes = [vp]
        L.vpin_kernel_launches.restype = u64
        L.vpin_kernel_launches.argtypes = [vp]
        L.vpin_stream.restype = vp
        L.vpin_stream.argtypes = [vp]
        L.vpin_last_phase_times.restype = C.c_uint32
        L.vpin_profile_read.restype = C.c_uint32
    return _lib


def _buf(b):
    """bytes / bytearray / numpy array -> c pointer (keeps a reference alive in the caller)."""
    if b is None:
        return None
    if isinstance(b, np.ndarray):
        return b.ctypes.data_as(C.c_void_p)
    if isinstance(b, (bytes, bytearray)):
        return C.cast(C.c_char_p(bytes(b)) if isinstance(b, bytes) else (C.c_char * len(b)).from_buffer(b), C.c_void_p)
    raise TypeError(type(b))


class Context:
    """One per host thread (the reference API is single-threaded: &mut Transcript, &mut RandomTape)."""

    def __init__(self, device=0, high_priority=False, background=False):
        """high_priority: the device's most urgent stream priority (the proof on the critical path); background: the lowest
        priority and one resident MSM block per SM (work nothing waits for, see SNARK.encode_commit)"""
        self._h = C.c_void_p()
        self.device = device
        st = lib().vpin_ctx_create_ex(C.c_int32(device), C.c_int32(-1 if background else (1 if high_priority else 0)), C.byref(self._h))
        if st != 0:
            raise VpinError(st, "vpin_ctx_create failed (no usable CUDA device?)")

    def init_distributed(self, rank, world, dist):
        """One process per GPU: rank 0 creates the NCCL id, `dist` (torch.distributed, any backend) ships it, every rank
        joins the communicator. Afterwards every rank must issue the same calls with the same inputs."""
        self.rank, self.world = rank, world
        if world == 1:
            return
        uid = exchange_unique_id(dist, rank)
        self.check(lib().vpin_ctx_init_distributed(self._h, C.c_int32(rank), C.c_int32(world), uid))

    def set_shard_sumcheck(self, on):
        """sharded sumcheck rounds of one proof across the ranks (1 / 0; -1 = the VPIN_SHARD_SUMCHECK environment default)"""
        self.check(lib().vpin_ctx_set_shard_sumcheck(self._h, C.c_int32(int(on))))

    def close(self):
        if self._h:
            lib().vpin_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, st):
        if st != 0:
            raise VpinError(st, lib().vpin_last_error(self._h).decode())

    @property
    def kernel_launches(self):
        return int(lib().vpin_kernel_launches(self._h))

    @property
    def stream(self):
        return int(lib().vpin_stream(self._h) or 0)

    def sync(self):
        self.check(lib().vpin_sync(self._h))

    def phase_times(self):
        names = (C.c_char_p * 32)()
        ms = (C.c_double * 32)()
        n = lib().vpin_last_phase_times(self._h, names, ms, 32)
        return {names[i].decode(): ms[i] for i in range(n)}

    # ---- kernel-level entry points (canonical 32-byte scalars in/out) ----
    def derive_gens(self, label, n):
        out = C.create_string_buffer(32 * (n + 1))
        self.check(lib().vpin_derive_gens(self._h, label, C.c_uint64(n), out))
        return out.raw

    def msm(self, label, scalars):
        out = C.create_string_buffer(32)
        self.check(lib().vpin_msm(self._h, label, scalars, C.c_uint64(len(scalars) // 32), out))
        return out.raw

    def hyrax_commit(self, label, Z, blinds=None):
        n = len(Z) // 32
        ell = n.bit_length() - 1
        L = 1 << (ell // 2)
        out = C.create_string_buffer(32 * L)
        self.check(lib().vpin_hyrax_commit(self._h, label, Z, C.c_uint64(n), blinds, out))
        return out.raw

    def eq_evals(self, r):
        ell = len(r) // 32
        out = C.create_string_buffer(32 << ell)
        self.check(lib().vpin_eq_evals(self._h, r, C.c_uint32(ell), out))
        return out.raw

    def sumcheck_cubic_round(self, A, B, Cc, D):
        out = C.create_string_buffer(96)
        self.check(lib().vpin_sumcheck_cubic_round(self._h, A, B, Cc, D, C.c_uint64(len(A) // 32), out))
        return out.raw

    def sumcheck_quad_round(self, A, B):
        out = C.create_string_buffer(64)
        self.check(lib().vpin_sumcheck_quad_round(self._h, A, B, C.c_uint64(len(A) // 32), out))
        return out.raw

    def sumcheck_cubic3_round(self, A, B, Cc):
        out = C.create_string_buffer(96)
        self.check(lib().vpin_sumcheck_cubic3_round(self._h, A, B, Cc, C.c_uint64(len(A) // 32), out))
        return out.raw

    def sumcheck_fused(self, degree, tables, r):
        """whole sumcheck through the fused bind+evaluate kernels -> (rounds x degree evals, finals), canonical bytes"""
        n = len(tables[0]) // 32
        rounds = n.bit_length() - 1
        evals = C.create_string_buffer(32 * rounds * degree)
        finals = C.create_string_buffer(32 * len(tables))
        t = list(tables) + [None] * (4 - len(tables))
        self.check(lib().vpin_sumcheck_fused(self._h, C.c_uint32(degree), t[0], t[1], t[2], t[3], C.c_uint64(n), r, evals, finals))
        return evals.raw, finals.raw

    def spark_timestamps(self, addrs, N, M):
        """AddrTimestamps::new for one side: addrs = three uint32 numpy arrays -> (addr[3N], read_ts[3N], audit_ts[M])"""
        a = [np.ascontiguousarray(x, dtype=np.uint32) for x in addrs]
        out_addr, out_ts, out_audit = np.empty(3 * N, np.uint32), np.empty(3 * N, np.uint32), np.empty(M, np.uint32)
        ptr = lambda x: x.ctypes.data_as(C.c_void_p) if len(x) else None
        self.check(lib().vpin_spark_timestamps(self._h, ptr(a[0]), C.c_uint64(len(a[0])), ptr(a[1]), C.c_uint64(len(a[1])), ptr(a[2]),
                                               C.c_uint64(len(a[2])), C.c_uint64(N), C.c_uint64(M), ptr(out_addr), ptr(out_ts), ptr(out_audit)))
        return out_addr, out_ts, out_audit

    def bind_top(self, Z, r):
        buf = C.create_string_buffer(Z, len(Z))
        self.check(lib().vpin_bind_top(self._h, buf, C.c_uint64(len(Z) // 32), r))
        return buf.raw[: len(Z) // 2]

    def bound(self, Z, Lvec):
        n = len(Z) // 32
        ell = n.bit_length() - 1
        R = 1 << (ell - ell // 2)
        out = C.create_string_buffer(32 * R)
        self.check(lib().vpin_bound(self._h, Z, C.c_uint64(n), Lvec, out))
        return out.raw

    def product_tree(self, leaves):
        n = len(leaves) // 32
        out = C.create_string_buffer(32 * (2 * n - 2))
        self.check(lib().vpin_product_tree(self._h, leaves, C.c_uint64(n), out))
        return out.raw

    def hash_layer(self, addr, vals, ts, gamma, tau):
        n = len(addr)
        rd, wr = C.create_string_buffer(32 * n), C.create_string_buffer(32 * n)
        self.check(lib().vpin_hash_layer(self._h, (C.c_uint32 * n)(*addr), vals, (C.c_uint32 * n)(*ts), C.c_uint64(n), gamma, tau, rd, wr))
        return rd.raw, wr.raw

    def deref_gather(self, addr, mem):
        n = len(addr)
        out = C.create_string_buffer(32 * n)
        self.check(lib().vpin_deref_gather(self._h, (C.c_uint32 * n)(*addr), C.c_uint64(n), mem, C.c_uint64(len(mem) // 32), out))
        return out.raw

    def commitments_add(self, c1, c2):
        out = C.create_string_buffer(len(c1))
        self.check(lib().vpin_commitments_add(self._h, c1, c2, C.c_uint64(len(c1) // 32), out))
        return out.raw

    def profile_enable(self, on=True, min_units=-1.0):
        """per-kernel-class CUDA-event timing on the context stream (see vpin_profile_enable in the header)"""
        self.check(lib().vpin_profile_enable(self._h, C.c_int32(1 if on else 0), C.c_double(min_units)))

    def profile_read(self):
        """-> ({class: dict(ms, launches, units, bytes)}, msm_madds)"""
        cap = 32
        names = (C.c_char_p * cap)()
        ms = (C.c_double * cap)()
        launches = (C.c_uint64 * cap)()
        units = (C.c_double * cap)()
        nbytes = (C.c_double * cap)()
        madds = C.c_uint64()
        n = lib().vpin_profile_read(self._h, names, ms, launches, units, nbytes, C.c_uint32(cap), C.byref(madds))
        out = {names[i].decode(): dict(ms=ms[i], launches=int(launches[i]), units=units[i], bytes=nbytes[i]) for i in range(n)}
        return out, int(madds.value)

    def imad_peak(self):
        v = C.c_double()
        self.check(lib().vpin_imad_peak(self._h, C.byref(v)))
        return v.value

    def imad_peak_forms(self):
        """(plain IMAD.WIDE product + ALU combine, single-instruction IMAD.WIDE multiply-accumulate) in MAC/s"""
        v = (C.c_double * 2)()
        self.check(lib().vpin_imad_peak_forms(self._h, v))
        return v[0], v[1]


def exchange_unique_id(dist, rank, make_id=None):
    """rank 0 makes the 128-byte NCCL unique id (vpin_nccl_unique_id), everyone receives it through `dist`."""
    box = [None]
    if rank == 0:
        if make_id is None:
            buf = C.create_string_buffer(128)
            st = lib().vpin_nccl_unique_id(buf)
            if st != 0:
                raise VpinError(st, "vpin_nccl_unique_id failed (libnccl not loadable?)")
            box[0] = buf.raw
        else:
            box[0] = make_id()
    dist.broadcast_object_list(box, src=0)
    assert isinstance(box[0], bytes) and len(box[0]) == 128
    return box[0]


def shard_rows(rows, rank, world):
    """(r0, r1, sharded): the Hyrax rows rank `rank` of `world` commits to (vpin_shard_rows)"""
    r0, r1, s = C.c_uint64(), C.c_uint64(), C.c_int32()
    lib().vpin_shard_rows(C.c_uint64(rows), C.c_int32(rank), C.c_int32(world), C.byref(r0), C.byref(r1), C.byref(s))
    return r0.value, r1.value, bool(s.value)


class SNARKGens:
    """SNARKGens::new(num_cons, num_vars, num_inputs, num_nz_entries)  (Spartan/src/lib.rs:305)"""

    def __init__(self, ctx, num_cons, num_vars, num_inputs, num_nz_entries):
        self.ctx = ctx
        self.params = (num_cons, num_vars, num_inputs, num_nz_entries)
        self._h = C.c_void_p()
        ctx.check(lib().vpin_gens_create(ctx._h, C.c_uint64(num_cons), C.c_uint64(num_vars), C.c_uint64(num_inputs),
                                         C.c_uint64(num_nz_entries), C.byref(self._h)))
        L, R = C.c_uint64(), C.c_uint64()
        lib().vpin_gens_witness_grid(self._h, C.byref(L), C.byref(R))
        self.L, self.R = L.value, R.value

    new = classmethod(lambda cls, *a: cls(*a))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().vpin_gens_destroy(self._h)
            self._h = None


class Instance:
    """Instance::new(num_cons, num_vars, num_inputs, &A, &B, &C)  (Spartan/src/lib.rs:138-244).
    A, B, C: numpy arrays of COO_DTYPE (row, col, 32-byte LE canonical value)."""

    def __init__(self, ctx, num_cons, num_vars, num_inputs, A, B, Cm, _handle=None):
        self.ctx = ctx
        self._h = C.c_void_p()
        if _handle is not None:
            self._h = _handle
        else:
            A, B, Cm = (np.ascontiguousarray(x, dtype=COO_DTYPE) for x in (A, B, Cm))
            ctx.check(lib().vpin_instance_create(ctx._h, C.c_uint64(num_cons), C.c_uint64(num_vars), C.c_uint64(num_inputs),
                                                 _buf(A), C.c_uint64(len(A)), _buf(B), C.c_uint64(len(B)), _buf(Cm),
                                                 C.c_uint64(len(Cm)), C.byref(self._h)))
        nc, nv, ni = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib().vpin_instance_dims(self._h, C.byref(nc), C.byref(nv), C.byref(ni))
        self.num_cons, self.num_vars, self.num_inputs = nc.value, nv.value, ni.value

    def is_sat(self, vars_bytes, inputs_bytes):
        sat = C.c_int32()
        self.ctx.check(lib().vpin_instance_is_sat(self.ctx._h, self._h, vars_bytes, C.c_uint64(len(vars_bytes) // 32), inputs_bytes,
                                                  C.c_uint64(len(inputs_bytes) // 32), C.byref(sat)))
        return bool(sat.value)

    def spmv_abc(self, z):
        n = 32 * self.num_cons
        outs = [C.create_string_buffer(n) for _ in range(3)]
        self.ctx.check(lib().vpin_spmv_abc(self.ctx._h, self._h, z, *outs))
        return tuple(o.raw for o in outs)

    def spmv_t_abc(self, x):
        n = 32 * 2 * self.num_vars
        outs = [C.create_string_buffer(n) for _ in range(3)]
        self.ctx.check(lib().vpin_spmv_t_abc(self.ctx._h, self._h, x, *outs))
        return tuple(o.raw for o in outs)

    def export_coo(self, num_vars_unpadded):
        """(A, B, C) as numpy COO_DTYPE arrays in Instance::new's input format"""
        nnz = (C.c_uint64 * 3)()
        lib().vpin_instance_nnz(self._h, nnz)
        out = []
        for k in range(3):
            a = np.zeros(nnz[k], COO_DTYPE)
            self.ctx.check(lib().vpin_instance_export_coo(self.ctx._h, self._h, C.c_uint64(num_vars_unpadded), C.c_int32(k), _buf(a)))
            out.append(a)
        return tuple(out)

    def pad(self, assignment):
        """Assignment::pad (Spartan/src/lib.rs:107-120)"""
        return assignment + bytes(32 * self.num_vars - len(assignment))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().vpin_instance_destroy(self._h)
            self._h = None


def point_mult(ctx, weights, px, py):
    """point_mult(network) of vPIN_proof_generation/src/point_mult.rs:7-664 with the JSON inputs passed in.
    Returns (dims, inst, vars_para, vars_input, vars, inputs) — dims = (num_cons, num_vars, num_inputs, num_non_zero_entries)."""
    m = len(weights)
    dims = (C.c_uint64 * 4)()
    lib().vpin_point_mult_dims(C.c_uint64(m), dims)
    nv = dims[1]
    w = (C.c_uint64 * (2 * m))(*[x for ww in weights for x in (ww & (2**64 - 1), ww >> 64)])
    bufs = [C.create_string_buffer(32 * nv) for _ in range(3)]
    inputs = C.create_string_buffer(32)
    h = C.c_void_p()
    ctx.check(lib().vpin_build_point_mult(ctx._h, C.c_uint64(m), w, px, py, C.byref(h), dims, bufs[0], bufs[1], bufs[2], inputs))
    inst = Instance(ctx, 0, 0, 0, None, None, None, _handle=h)
    return tuple(dims), inst, bufs[0].raw, bufs[1].raw, bufs[2].raw, inputs.raw


def point_mult_device(ctx, weights, px, py):
    """point_mult with the three assignments left in HBM (vpin_build_point_mult_device): returns
    (dims, inst, d_vars_para, d_vars_input, d_vars, inputs, padded) — torch uint8 device tensors of padded x 32 bytes in
    Montgomery form, ready for dev_poly_commit* / DeviceWitness."""
    import torch
    m = len(weights)
    dims = (C.c_uint64 * 4)()
    lib().vpin_point_mult_dims(C.c_uint64(m), dims)
    padded = 1 << max(1, (int(dims[1]) - 1).bit_length())
    dev = torch.device("cuda", ctx.device)
    bufs = [torch.empty(32 * padded, dtype=torch.uint8, device=dev) for _ in range(3)]
    torch.cuda.synchronize(dev)
    w = (C.c_uint64 * (2 * m))(*[x for ww in weights for x in (ww & (2**64 - 1), ww >> 64)])
    inputs = C.create_string_buffer(32)
    h = C.c_void_p()
    ctx.check(lib().vpin_build_point_mult_device(ctx._h, C.c_uint64(m), w, px, py, C.byref(h), dims, _dp(bufs[0]), _dp(bufs[1]), _dp(bufs[2]),
                                                 C.c_uint64(padded), inputs))
    inst = Instance(ctx, 0, 0, 0, None, None, None, _handle=h)
    return tuple(dims), inst, bufs[0], bufs[1], bufs[2], inputs.raw, padded


def point_addition(ctx, px, py, rx, ry, rz):
    """point_addition(network) of vPIN_proof_generation/src/point_addition.rs:5-326."""
    n = len(rz)
    dims = (C.c_uint64 * 4)()
    lib().vpin_point_add_dims(C.c_uint64(n), dims)
    nv = dims[1]
    bufs = [C.create_string_buffer(32 * nv) for _ in range(3)]
    h = C.c_void_p()
    ctx.check(lib().vpin_build_point_add(ctx._h, C.c_uint64(n), px, py, rx, ry, (C.c_int64 * n)(*rz), C.byref(h), dims, bufs[0], bufs[1],
                                         bufs[2]))
    inst = Instance(ctx, 0, 0, 0, None, None, None, _handle=h)
    return tuple(dims), inst, bufs[0].raw, bufs[1].raw, bufs[2].raw, b""


class RandomTape:
    """RandomTape::new(name) (Spartan/src/random.rs:14-21). init_randomness32=None: the library draws the OsRng scalar itself
    (production); 32 canonical bytes reproduce a run (tests only - a seed must never serve two different witnesses)."""

    def __init__(self, name, init_randomness32=None):
        self.state = C.create_string_buffer(256)
        st = lib().vpin_tape_init(self.state, name, C.c_uint64(len(name)), init_randomness32)
        if st != 0:
            raise VpinError(st, "vpin_tape_init")


class SNARK:
    @staticmethod
    def encode(inst, gens, ctx=None):
        """SNARK::encode(&inst, &gens) -> (ComputationCommitment as bincode bytes, ComputationDecommitment handle).
        ctx: the context (stream, scratch) the call runs on, the instance's own by default. Encoding does not depend on the
        witness, so a driver may run it on a SECOND context from a second host thread while the first one commits to the
        assignments (bench.py): the decommitment can be handed to a proof on any context of the device."""
        ctx = ctx or inst.ctx
        cap = 64 + 32 * (1 << 16) * 2
        out = getattr(ctx, "_comm_buf", None)  # one 4 MB output buffer per context, reused (no mmap / munmap per call)
        if out is None:
            out = ctx._comm_buf = C.create_string_buffer(cap)
        n = C.c_uint64()
        d = C.c_void_p()
        st = lib().vpin_encode(ctx._h, inst._h, gens._h, out, C.c_uint64(cap), C.byref(n), C.byref(d))
        ctx.check(st)
        return C.string_at(out, n.value), Decommitment(d)


def _comm_buffer(ctx):
    cap = 64 + 32 * (1 << 16) * 2
    out = getattr(ctx, "_comm_buf", None)  # one 4 MB output buffer per context, reused (no mmap / munmap per call)
    if out is None:
        out = ctx._comm_buf = C.create_string_buffer(cap)
    return out, cap


def encode_tables(inst, gens, ctx=None):
    """First half of SNARK::encode: the dense representation my_lib_prove reads (no commitment) -> Decommitment handle"""
    ctx = ctx or inst.ctx
    d = C.c_void_p()
    ctx.check(lib().vpin_encode_tables(ctx._h, inst._h, gens._h, C.byref(d)))
    return Decommitment(d)


def encode_commit(decomm, gens, ctx):
    """Second half: the ComputationCommitment (bincode bytes) of a Decommitment. Nothing in my_lib_prove depends on it: run it on
    a background context (Context(device, background=True)) from a helper thread while the proof is under way."""
    out, cap = _comm_buffer(ctx)
    n = C.c_uint64()
    ctx.check(lib().vpin_encode_commit(ctx._h, decomm._h, gens._h, out, C.c_uint64(cap), C.byref(n)))
    return C.string_at(out, n.value)


class Decommitment:
    def __init__(self, h):
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            lib().vpin_decomm_destroy(self._h)
            self._h = None


def dense_mlpoly_commit(ctx, gens, Z, tape, n=None):
    """DensePolynomial::new(Z).commit(&gens.gens_r1cs_sat.gens_pc, Some(&mut tape)) -> (PolyCommitment, blinds).
    Z: bytes, or a host pointer (c_void_p) together with n."""
    out = C.create_string_buffer(32 * gens.L)
    blinds = C.create_string_buffer(32 * gens.L)
    n = len(Z) // 32 if n is None else n
    ctx.check(lib().vpin_poly_commit(ctx._h, gens._h, Z, C.c_uint64(n), tape.state if tape else None, out, blinds))
    return out.raw, blinds.raw


def my_dense_mlpoly_commit(ctx, gens, Z, blind_1, blind_2, n=None):
    """vPIN_proof_generation/src/commit_test.rs:27-57"""
    out = C.create_string_buffer(32 * gens.L)
    blinds = C.create_string_buffer(32 * gens.L)
    n = len(Z) // 32 if n is None else n
    ctx.check(lib().vpin_poly_commit_with_blinds(ctx._h, gens._h, Z, C.c_uint64(n), blind_1, blind_2, C.c_uint64(gens.L), out, blinds))
    return out.raw, blinds.raw


class Witness:
    """(vars, poly_vars, comm_vars, blinds_vars) resident in HBM."""

    def __init__(self, ctx, gens, vars_bytes, comm_vars, blinds_vars):
        self._h = C.c_void_p()
        ctx.check(lib().vpin_witness_upload(ctx._h, gens._h, vars_bytes, C.c_uint64(len(vars_bytes) // 32), comm_vars, blinds_vars,
                                            C.c_uint64(gens.L), C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().vpin_witness_destroy(self._h)
            self._h = None


_PROOF_CAP = 8 << 20


def _proof_buffer(ctx):
    """one 8 MB output buffer per context, reused (allocating and zeroing a fresh one costs ~1 ms per proof)"""
    buf = getattr(ctx, "_proof_buf", None)
    if buf is None:
        buf = ctx._proof_buf = C.create_string_buffer(_PROOF_CAP)
    return buf


def my_lib_prove(inst, decomm, vars_bytes, inputs_bytes, gens, transcript_label, comm_vars, blinds_vars, tape_seed, n=None, ctx=None):
    """my_lib_prove (vPIN_proof_generation/src/commit_test.rs:59-133): host buffers in, bincode(SNARK) out.
    `vars_bytes` doubles as poly_vars (DensePolynomial::new(padded_vars.assignment)). ctx: the context the proof runs on (the
    instance's own by default; an instance or decommitment made on another context of the device may be proved on any)."""
    ctx = ctx or inst.ctx
    out = _proof_buffer(ctx)
    n_vars = len(vars_bytes) // 32 if n is None else n
    n = C.c_uint64()
    ctx.check(lib().vpin_prove(ctx._h, inst._h, decomm._h, vars_bytes, C.c_uint64(n_vars), inputs_bytes,
                               C.c_uint64(len(inputs_bytes) // 32), gens._h, transcript_label, C.c_uint64(len(transcript_label)),
                               comm_vars, blinds_vars, C.c_uint64(gens.L), tape_seed, out, C.c_uint64(_PROOF_CAP), C.byref(n)))
    return C.string_at(out, n.value)


def my_lib_prove_resident(inst, decomm, witness, inputs_bytes, gens, transcript_label, tape_seed):
    ctx = inst.ctx
    out = _proof_buffer(ctx)
    n = C.c_uint64()
    ctx.check(lib().vpin_prove_resident(ctx._h, inst._h, decomm._h, witness._h, inputs_bytes, C.c_uint64(len(inputs_bytes) // 32), gens._h,
                                        transcript_label, C.c_uint64(len(transcript_label)), tape_seed, out, C.c_uint64(_PROOF_CAP),
                                        C.byref(n)))
    return C.string_at(out, n.value)


def prove_flow(ctx, dims, inst, vars_para, vars_input, vars_, inputs, seed_q, seed_p, label=b"snark_example"):
    """The driver sequence of vPIN_proof_generation/src/proof_point_add.rs:39-98 (== proof_point_mult.rs):
    gens -> encode -> commit(para) -> commit(input) -> my_dense_mlpoly_commit -> row-wise combine -> my_lib_prove.
    Returns a dict with the proof, the computation commitment and the three witness commitments."""
    num_cons, num_vars, num_inputs, nnz = dims
    gens = SNARKGens(ctx, num_cons, num_vars, num_inputs, nnz)
    comm, decomm = SNARK.encode(inst, gens)
    tape = RandomTape(b"\x02", seed_q)
    p_para, p_input, p_vars = inst.pad(vars_para), inst.pad(vars_input), inst.pad(vars_)
    c_para, b_para = dense_mlpoly_commit(ctx, gens, p_para, tape)
    c_input, b_input = dense_mlpoly_commit(ctx, gens, p_input, tape)
    c_vars, b_vars = my_dense_mlpoly_commit(ctx, gens, p_vars, b_para, b_input)
    combined = ctx.commitments_add(c_para, c_input)
    if combined[:32] != c_vars[:32]:
        raise VpinError(8, "commitment homomorphism check failed (proof_point_add.rs:69-73)")
    proof = my_lib_prove(inst, decomm, p_vars, inputs, gens, label, combined, b_vars, seed_p)
    return dict(proof=proof, comm=comm, comm_vars_para=c_para, comm_vars_input=c_input, comm_vars=c_vars, gens=gens, decomm=decomm,
                padded_vars=p_vars, blinds_vars=b_vars, combined=combined)


# ---- HBM-resident forms (device pointers, Montgomery scalars): what bench.py times as `value` ----
def _dp(x):
    """torch tensor or int -> device pointer"""
    return C.c_void_p(x if isinstance(x, int) else x.data_ptr())


def dev_to_mont(ctx, d_in, n, d_out):
    ctx.check(lib().vpin_dev_to_mont(ctx._h, _dp(d_in), C.c_uint64(n), _dp(d_out)))


def dev_from_mont(ctx, d_in, n, d_out):
    ctx.check(lib().vpin_dev_from_mont(ctx._h, _dp(d_in), C.c_uint64(n), _dp(d_out)))


def dev_poly_commit(ctx, gens, d_Z, n, tape, d_points_out, d_blinds_out):
    ctx.check(lib().vpin_dev_poly_commit(ctx._h, gens._h, _dp(d_Z), C.c_uint64(n), tape.state if tape else None, _dp(d_points_out),
                                         _dp(d_blinds_out)))


def dev_poly_commit_with_blinds(ctx, gens, d_Z, n, d_b1, d_b2, d_points_out, d_blinds_out):
    ctx.check(lib().vpin_dev_poly_commit_with_blinds(ctx._h, gens._h, _dp(d_Z), C.c_uint64(n), _dp(d_b1), _dp(d_b2), _dp(d_points_out),
                                                     _dp(d_blinds_out)))


def dev_commitments_add(ctx, d_c1, d_c2, L, d_out):
    ctx.check(lib().vpin_dev_commitments_add(ctx._h, _dp(d_c1), _dp(d_c2), C.c_uint64(L), _dp(d_out)))


def prove_flow_resident(ctx, weights, px, py, seed_q, seed_p, label=b"snark_example", timings=None):
    """The point-mult driver sequence (vPIN_proof_generation/src/proof_point_mult.rs:24-98) with NOTHING large crossing PCIe:
    the R1CS instance and the three assignments are expanded on the device from the JSON-level inputs
    (vpin_build_point_mult_device), committed where they lie (vpin_dev_poly_commit*) and proved resident
    (vpin_prove_resident). Host traffic: the weights and point coordinates up (80 B per multiplication), the commitments
    (32 B per Hyrax row) and the proof down. `timings` (a dict) receives the wall time of each call in seconds."""
    import time
    import torch
    T = [time.time()]

    def tick(name):
        ctx.sync()
        T.append(time.time())
        if timings is not None:
            timings[name] = timings.get(name, 0.0) + T[-1] - T[-2]

    dims, inst, d_para, d_input, d_vars, inputs, n = point_mult_device(ctx, weights, px, py)
    tick("point_mult (device build + Instance::new)")
    gens = SNARKGens(ctx, *dims)
    tick("SNARKGens::new")
    comm, decomm = SNARK.encode(inst, gens)
    tick("SNARK::encode")
    dev = d_vars.device
    L = gens.L
    pts = [torch.empty(32 * L, dtype=torch.uint8, device=dev) for _ in range(4)]
    blinds = [torch.empty(32 * L, dtype=torch.uint8, device=dev) for _ in range(3)]
    tape = RandomTape(b"\x02", seed_q)
    dev_poly_commit(ctx, gens, d_para, n, tape, pts[0], blinds[0])
    dev_poly_commit(ctx, gens, d_input, n, tape, pts[1], blinds[1])
    dev_poly_commit_with_blinds(ctx, gens, d_vars, n, blinds[0], blinds[1], pts[2], blinds[2])
    dev_commitments_add(ctx, pts[0], pts[1], L, pts[3])
    tick("3 commits + combine")
    wit = DeviceWitness(ctx, gens, d_vars, n, pts[3], blinds[2])
    proof = my_lib_prove_resident(inst, decomm, wit, inputs, gens, label, seed_p)
    tick("my_lib_prove")
    ctx.sync()
    c_para, c_input, c_vars = (bytes(t.cpu().numpy()) for t in pts[:3])
    tick("commitments to the host")
    return dict(proof=proof, comm=comm, comm_vars_para=c_para, comm_vars_input=c_input, comm_vars=c_vars, dims=dims, inputs=inputs)


class DeviceWitness(Witness):
    """Witness built from device-resident (vars, comm_vars, blinds_vars)"""

    def __init__(self, ctx, gens, d_vars, n_vars, d_comm, d_blinds):
        self._h = C.c_void_p()
        ctx.check(lib().vpin_witness_from_device(ctx._h, gens._h, _dp(d_vars), C.c_uint64(n_vars), _dp(d_comm), _dp(d_blinds),
                                                 C.c_uint64(gens.L), C.byref(self._h)))
