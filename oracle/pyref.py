"""TEST INFRASTRUCTURE ONLY — pure-Python big-int restatement of the primitives under the
Spartan prover path (tier-0 oracle, "obviously correct, slow").

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
It pins the C++ oracle (oracle/*.hpp) and, through it, the CUDA product path.

What is restated and where it comes from:
  * F_l   scalar field, Montgomery R = 2^256:  Spartan/src/scalar/ristretto255.rs:200-329 (constants),
          :398-473 (from_bytes / to_bytes / from_bytes_wide).
  * ristretto255 over Edwards25519: NOT in /root/reference — dependency curve25519-dalek 3.2.0
          (Spartan/Cargo.toml:14).  Restated from RFC 9496 (DECODE / ENCODE / MAP / from_uniform_bytes),
          pinned against RFC 9496 vectors and libsodium 1.0.20 in tests/test_oracle_primitives.py.
  * Merlin transcripts over STROBE-128 / Keccak-f[1600]: NOT in /root/reference — dependency
          merlin 3.0.0.  Restated from the STROBE v1.0.2 + Merlin specs; pinned by the Merlin
          crate's published test vector (tests/test_oracle_primitives.py).
  * Generator derivation: Spartan/src/commitments.rs:20-38 (SHAKE256(label || basepoint) XOF).
  * Transcript helpers: Spartan/src/transcript.rs:19-43, Spartan/src/random.rs:14-30.
"""
import hashlib

# ----------------------------------------------------------------------------- F_l
L = 2**252 + 27742317777372353535851937790883648493
R_MONT = (1 << 256) % L


def fl_to_mont_bytes(x):
    """Scalar serde = raw Montgomery limbs (ristretto255.rs:199-200)."""
    return ((x * R_MONT) % L).to_bytes(32, "little")


def fl_from_mont_bytes(b):
    return (int.from_bytes(b, "little") * pow(R_MONT, -1, L)) % L


def fl_from_bytes_wide(b64):
    """ristretto255.rs:442-473: 512-bit LE integer reduced mod l."""
    return int.from_bytes(b64, "little") % L


# ----------------------------------------------------------------------------- F_p, Edwards25519
P = 2**255 - 19
D = (-121665 * pow(121666, P - 2, P)) % P
SQRT_M1 = pow(2, (P - 1) // 4, P)
# RFC 9496 section 4.1 constants (recomputed in tests, sign pinned by the RFC vectors)
SQRT_AD_MINUS_ONE = 25063068953384623474111414158702152701244531502492656460079210482610430750235
INVSQRT_A_MINUS_D = 54469307008909316920995813868745141605393597292927456921205312896311721017578
ONE_MINUS_D_SQ = (1 - D * D) % P
D_MINUS_ONE_SQ = ((D - 1) ** 2) % P


def _is_neg(x):
    return (x % P) & 1


def _abs(x):
    x %= P
    return P - x if x & 1 else x


def sqrt_ratio_m1(u, v):
    """RFC 9496 4.2 SQRT_RATIO_M1."""
    u %= P
    v %= P
    r = (u * pow(v, 3, P) * pow(u * pow(v, 7, P), (P - 5) // 8, P)) % P
    check = (v * r * r) % P
    correct = check == u
    flipped = check == (-u) % P
    flipped_i = check == (-u * SQRT_M1) % P
    if flipped or flipped_i:
        r = (r * SQRT_M1) % P
    return (correct or flipped), _abs(r)


IDENT = (0, 1, 1, 0)


def pt_add(p, q):
    """Extended twisted Edwards, a = -1 (add-2008-hwcd-3)."""
    x1, y1, z1, t1 = p
    x2, y2, z2, t2 = q
    a = ((y1 - x1) * (y2 - x2)) % P
    b = ((y1 + x1) * (y2 + x2)) % P
    c = (t1 * 2 * D * t2) % P
    d = (z1 * 2 * z2) % P
    e, f, g, h = b - a, d - c, d + c, b + a
    return ((e * f) % P, (g * h) % P, (f * g) % P, (e * h) % P)


def pt_neg(p):
    x, y, z, t = p
    return ((-x) % P, y, z, (-t) % P)


def pt_mul(k, p):
    k %= L
    acc = IDENT
    while k:
        if k & 1:
            acc = pt_add(acc, p)
        p = pt_add(p, p)
        k >>= 1
    return acc


def msm(scalars, points):
    acc = IDENT
    for s, p in zip(scalars, points):
        acc = pt_add(acc, pt_mul(s, p))
    return acc


def ristretto_decode(b):
    """RFC 9496 4.3.1 DECODE; returns None on failure."""
    s = int.from_bytes(b, "little")
    if s >= P or (s & 1):
        return None
    ss = (s * s) % P
    u1 = (1 - ss) % P
    u2 = (1 + ss) % P
    u2_sqr = (u2 * u2) % P
    v = (-(D * u1 * u1) - u2_sqr) % P
    ok, invsqrt = sqrt_ratio_m1(1, (v * u2_sqr) % P)
    den_x = (invsqrt * u2) % P
    den_y = (invsqrt * den_x * v) % P
    x = _abs(2 * s * den_x)
    y = (u1 * den_y) % P
    t = (x * y) % P
    if (not ok) or _is_neg(t) or y == 0:
        return None
    return (x, y, 1, t)


def ristretto_encode(p):
    """RFC 9496 4.3.2 ENCODE."""
    x0, y0, z0, t0 = p
    u1 = ((z0 + y0) * (z0 - y0)) % P
    u2 = (x0 * y0) % P
    _, invsqrt = sqrt_ratio_m1(1, (u1 * u2 * u2) % P)
    den1 = (invsqrt * u1) % P
    den2 = (invsqrt * u2) % P
    z_inv = (den1 * den2 * t0) % P
    ix0 = (x0 * SQRT_M1) % P
    iy0 = (y0 * SQRT_M1) % P
    enchanted = (den1 * INVSQRT_A_MINUS_D) % P
    rotate = _is_neg(t0 * z_inv)
    if rotate:
        x, y, den_inv = iy0, ix0, enchanted
    else:
        x, y, den_inv = x0, y0, den2
    if _is_neg(x * z_inv):
        y = (-y) % P
    s = _abs(den_inv * (z0 - y))
    return s.to_bytes(32, "little")


def _elligator(t):
    """RFC 9496 4.3.4 MAP."""
    r = (SQRT_M1 * t * t) % P
    u = ((r + 1) * ONE_MINUS_D_SQ) % P
    v = ((-1 - r * D) * (r + D)) % P
    was_square, s = sqrt_ratio_m1(u, v)
    s_prime = (-_abs(s * t)) % P
    s = s if was_square else s_prime
    c = (P - 1) if was_square else r
    n = (c * (r - 1) * D_MINUS_ONE_SQ - v) % P
    w0 = (2 * s * v) % P
    w1 = (n * SQRT_AD_MINUS_ONE) % P
    w2 = (1 - s * s) % P
    w3 = (1 + s * s) % P
    return ((w0 * w3) % P, (w2 * w1) % P, (w1 * w3) % P, (w0 * w2) % P)


def from_uniform_bytes(b64):
    """dalek RistrettoPoint::from_uniform_bytes == RFC 9496 one-way map (mask bit 255 of each half)."""
    t1 = int.from_bytes(b64[:32], "little") & ((1 << 255) - 1)
    t2 = int.from_bytes(b64[32:], "little") & ((1 << 255) - 1)
    return pt_add(_elligator(t1 % P), _elligator(t2 % P))


BASEPOINT_COMPRESSED = bytes.fromhex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76")


def derive_gens(label, n):
    """commitments.rs:20-38: n+1 points from one SHAKE256 stream; last one is h."""
    stream = hashlib.shake_256(label + BASEPOINT_COMPRESSED).digest(64 * (n + 1))
    pts = [from_uniform_bytes(stream[64 * i:64 * i + 64]) for i in range(n + 1)]
    return pts[:n], pts[n]


# ----------------------------------------------------------------------------- Keccak / STROBE / Merlin
_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B,
       0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088,
       0x0000000080008009, 0x000000008000000A, 0x000000008000808B, 0x800000000000008B, 0x8000000000008089,
       0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
       0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
_M64 = (1 << 64) - 1


def _rol(x, n):
    n %= 64
    return ((x << n) | (x >> (64 - n))) & _M64 if n else x


def keccak_f1600(state_bytes):
    a = [[int.from_bytes(state_bytes[8 * (x + 5 * y):8 * (x + 5 * y) + 8], "little") for y in range(5)] for x in range(5)]
    for rnd in range(24):
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                b[y][(2 * x + 3 * y) % 5] = _rol(a[x][y], _ROT[x][y])
        a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        a[0][0] ^= _RC[rnd]
    out = bytearray(200)
    for x in range(5):
        for y in range(5):
            out[8 * (x + 5 * y):8 * (x + 5 * y) + 8] = a[x][y].to_bytes(8, "little")
    return out


class Strobe128:
    R = 166
    FLAG_I, FLAG_A, FLAG_C, FLAG_T, FLAG_M, FLAG_K = 1, 2, 4, 8, 16, 32

    def __init__(self, protocol_label):
        st = bytearray(200)
        st[0:6] = bytes([1, self.R + 2, 1, 0, 1, 96])
        st[6:18] = b"STROBEv1.0.2"
        self.st = keccak_f1600(st)
        self.pos = 0
        self.pos_begin = 0
        self.cur_flags = 0
        self.meta_ad(protocol_label, False)

    def _run_f(self):
        self.st[self.pos] ^= self.pos_begin
        self.st[self.pos + 1] ^= 0x04
        self.st[self.R + 1] ^= 0x80
        self.st = keccak_f1600(self.st)
        self.pos = 0
        self.pos_begin = 0

    def _absorb(self, data):
        for byte in data:
            self.st[self.pos] ^= byte
            self.pos += 1
            if self.pos == self.R:
                self._run_f()

    def _squeeze(self, n):
        out = bytearray(n)
        for i in range(n):
            out[i] = self.st[self.pos]
            self.st[self.pos] = 0
            self.pos += 1
            if self.pos == self.R:
                self._run_f()
        return bytes(out)

    def _begin_op(self, flags, more):
        if more:
            assert self.cur_flags == flags
            return
        assert not (flags & self.FLAG_T)
        old_begin = self.pos_begin
        self.pos_begin = self.pos + 1
        self.cur_flags = flags
        self._absorb(bytes([old_begin, flags]))
        force_f = flags & (self.FLAG_C | self.FLAG_K)
        if force_f and self.pos != 0:
            self._run_f()

    def meta_ad(self, data, more):
        self._begin_op(self.FLAG_M | self.FLAG_A, more)
        self._absorb(data)

    def ad(self, data, more):
        self._begin_op(self.FLAG_A, more)
        self._absorb(data)

    def prf(self, n, more):
        self._begin_op(self.FLAG_I | self.FLAG_A | self.FLAG_C, more)
        return self._squeeze(n)


class Transcript:
    """merlin 3.0.0 Transcript + Spartan's ProofTranscript helpers (transcript.rs:19-43)."""

    def __init__(self, label):
        self.s = Strobe128(b"Merlin v1.0")
        self.append_message(b"dom-sep", label)

    def append_message(self, label, msg):
        self.s.meta_ad(label, False)
        self.s.meta_ad(len(msg).to_bytes(4, "little"), True)
        self.s.ad(msg, False)

    def append_u64(self, label, x):
        self.append_message(label, x.to_bytes(8, "little"))

    def challenge_bytes(self, label, n):
        self.s.meta_ad(label, False)
        self.s.meta_ad(n.to_bytes(4, "little"), True)
        return self.s.prf(n, False)

    # Spartan layer
    def append_protocol_name(self, name):
        self.append_message(b"protocol-name", name)

    def append_scalar(self, label, x):
        self.append_message(label, (x % L).to_bytes(32, "little"))

    def append_point(self, label, comp32):
        self.append_message(label, comp32)

    def challenge_scalar(self, label):
        return fl_from_bytes_wide(self.challenge_bytes(label, 64))

    def challenge_vector(self, label, n):
        return [self.challenge_scalar(label) for _ in range(n)]


def random_tape(name, init_randomness):
    """random.rs:14-21 with the OsRng scalar injected."""
    t = Transcript(name)
    t.append_scalar(b"init_randomness", init_randomness)
    return t


# ----------------------------------------------------------------------------- tiny polynomial helpers (KATs)
def eq_evals(r):
    """dense_mlpoly.rs:78-94 — r[0] is the most significant index bit."""
    evals = [1]
    for rj in r:
        nxt = []
        for e in evals:
            hi = (e * rj) % L
            nxt.append((e - hi) % L)
            nxt.append(hi)
        evals = nxt
    return evals


def mle_evaluate(Z, r):
    """dense_mlpoly.rs:249-255."""
    chis = eq_evals(r)
    return sum(z * c for z, c in zip(Z, chis)) % L


def unipoly_from_evals(evals):
    """unipoly.rs:23-54."""
    inv = lambda x: pow(x, L - 2, L)
    if len(evals) == 3:
        c = evals[0]
        a = inv(2) * (evals[2] - 2 * evals[1] + c) % L
        b = (evals[1] - c - a) % L
        return [c, b, a]
    d = evals[0]
    a = inv(6) * (evals[3] - 3 * evals[2] + 3 * evals[1] - evals[0]) % L
    b = inv(2) * (2 * evals[0] - 5 * evals[1] + 4 * evals[2] - evals[3]) % L
    c = (evals[1] - d - a - b) % L
    return [d, c, b, a]
