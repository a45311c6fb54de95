"""Pins the CPU oracle (oracle/*.hpp through liboracle.so, and the tier-0 oracle/pyref.py) against
  * the reference's own known-answer tests (Spartan/src/scalar/ristretto255.rs:788-1214, unipoly.rs:127-181,
    dense_mlpoly.rs:447-466) transcribed in tests/golden/kat_vectors.json,
  * RFC 9496 ristretto255 vectors and the merlin crate's transcript vector (the dalek / merlin boundary that the
    reference's tests do not pin),
  * an independent implementation present in this image (libsodium via ctypes), when it can be found.
CPU only (no GPU marker)."""
import ctypes as C
import glob
import hashlib
import json
import os
import random
import sys

import pytest

import oracle_lib as O

sys.path.insert(0, O.ORACLE_DIR)
import pyref as P  # noqa: E402

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat_vectors.json")))
L = O.L_ORDER
R = O.R_MONT


def limbs(v):
    return sum(int(x, 16) << (64 * i) for i, x in enumerate(v))


def mont(x):
    return ((x * R) % L).to_bytes(32, "little")


def unmont(b):
    return (int.from_bytes(b, "little") * O.R_INV) % L


def binop(op, a, b=0):
    out = C.create_string_buffer(32)
    O.lib().orc_fl_binop_mont(C.c_int(op), mont(a), mont(b), out)
    return unmont(out.raw)


ADD, SUB, MUL, NEG, INV, SQR = range(6)


# ------------------------------------------------------------------------------------------------ reference field KATs
def test_field_constants():
    f = KAT["reference_field"]
    assert limbs(f["MODULUS_limbs"]) == L
    assert limbs(f["R_limbs"]) == R
    assert limbs(f["R2_limbs"]) == R * R % L
    assert limbs(f["R3_limbs"]) == R * R * R % L
    assert int(f["INV"], 16) == (-pow(L, -1, 1 << 64)) % (1 << 64)


def test_to_bytes_from_bytes():
    f = KAT["reference_field"]
    lib = O.lib()
    out = C.create_string_buffer(32)
    for value, expect in ((0, bytes(32)), (1, bytes([1] + [0] * 31)), (R, bytes(f["to_bytes_R2"])), (L - 1, bytes(f["to_bytes_minus_one"]))):
        # the reference's `R2` constant is the Montgomery form of R, i.e. the scalar whose canonical value is R
        lib.orc_fl_to_bytes(mont(value), out)
        assert out.raw == expect
        m = C.create_string_buffer(32)
        assert lib.orc_fl_from_bytes(expect, m) == 1
        assert m.raw == mont(value)
    for bad in f["invalid_from_bytes"]:
        assert lib.orc_fl_from_bytes(bytes(bad), C.create_string_buffer(32)) == 0


def test_from_bytes_wide():
    f = KAT["reference_field"]
    lib = O.lib()
    out = C.create_string_buffer(32)
    lib.orc_fl_from_bytes_wide(bytes(f["to_bytes_R2"]) + bytes(32), out)
    assert out.raw == mont(R)
    lib.orc_fl_from_bytes_wide(bytes(f["to_bytes_minus_one"]) + bytes(32), out)
    assert out.raw == mont(L - 1)
    lib.orc_fl_from_bytes_wide(bytes([0xFF] * 64), out)
    # Scalar::from_raw(x) = x * R2 (Montgomery form of x): the raw limbs in the test are the canonical value
    assert unmont(out.raw) == limbs(f["from_bytes_wide_ff64_raw_limbs"]) % L == (2**512 - 1) % L
    assert P.fl_from_bytes_wide(bytes([0xFF] * 64)) == (2**512 - 1) % L
    assert limbs(f["from_raw_all_ff_equals_raw"]) == (2**256 - 1) % L


def test_add_sub_neg_around_largest():
    f = KAT["reference_field"]
    largest = limbs(f["LARGEST_limbs"])
    assert largest == L - 1
    # the reference compares raw Montgomery limbs: LARGEST + LARGEST = l - 2 as limb patterns
    assert binop(ADD, largest, largest) == limbs(f["LARGEST_plus_LARGEST_limbs"])
    assert binop(ADD, largest, 1) == 0
    assert binop(NEG, largest) == 1 and binop(NEG, 0) == 0 and binop(NEG, 1) == largest
    assert binop(SUB, largest, largest) == 0
    assert binop(SUB, 0, largest) == 1


def test_mul_square_invert_walk():
    """the reference walks cur = LARGEST, LARGEST + LARGEST, ... for 100 steps (ristretto255.rs:1083-1184)"""
    cur = L - 1
    for _ in range(100):
        assert binop(MUL, cur, cur) == cur * cur % L
        assert binop(SQR, cur) == cur * cur % L
        cur = (cur + L - 1) % L
    tmp = R * R % L
    for _ in range(100):
        inv = binop(INV, tmp)
        assert inv * tmp % L == 1 and inv == pow(tmp, L - 2, L)
        tmp = (tmp + R * R) % L
    assert binop(INV, 1) == 1 and binop(INV, L - 1) == L - 1


def test_field_random_against_bigint():
    rng = random.Random(1)
    for _ in range(2000):
        a, b = rng.randrange(L), rng.randrange(L)
        assert binop(ADD, a, b) == (a + b) % L
        assert binop(SUB, a, b) == (a - b) % L
        assert binop(MUL, a, b) == a * b % L


# ------------------------------------------------------------------------------------------------ unipoly / MLE KATs
def test_unipoly_from_evals():
    for key in ("quad", "cubic"):
        k = KAT["reference_unipoly"][key]
        coeffs = P.unipoly_from_evals(k["evals"])
        assert coeffs == k["coeffs"]
        assert sum(c * k["at"] ** i for i, c in enumerate(coeffs)) % L == k["value"]


def test_mle_evaluation_and_eq_order():
    k = KAT["reference_mle"]
    assert P.mle_evaluate(k["Z"], k["r"]) == k["value"]
    # eq tables: C++ oracle == tier-0 oracle == naive definition with r[0] on the most significant index bit
    rng = random.Random(2)
    for ell in (0, 1, 2, 5, 9):
        r = [rng.randrange(L) for _ in range(ell)]
        naive = []
        for i in range(1 << ell):
            v = 1
            for j in range(ell):
                bit = (i >> (ell - 1 - j)) & 1
                v = v * (r[j] if bit else (1 - r[j])) % L
            naive.append(v)
        assert P.eq_evals(r) == naive
        assert O.eq_evals(r) == naive


# ------------------------------------------------------------------------------------------------ ristretto255 (dalek boundary)
def test_rfc9496_basepoint_multiples():
    vec = [bytes.fromhex(h) for h in KAT["rfc9496_basepoint_multiples"]]
    B = P.ristretto_decode(vec[1])
    acc = None
    for k, enc in enumerate(vec):
        got = P.ristretto_encode(acc) if acc is not None else bytes(32)
        assert got == enc, k
        acc = B if acc is None else P.pt_add(acc, B)
    # C++ oracle: k*B by repeated addition of compressed points
    cur = vec[1]
    out = C.create_string_buffer(32)
    for k in range(2, 16):
        assert O.lib().orc_pt_add(cur, vec[1], out) == 1
        cur = out.raw
        assert cur == vec[k]
    # and through the MSM entry point
    for k in (1, 2, 7, 15):
        assert O.msm([k], vec[1]) == vec[k]
    assert O.msm([3, 5], vec[2] + vec[1]) == vec[11]


def test_rfc9496_bad_encodings_rejected():
    for h in KAT["rfc9496_bad_encodings"]:
        assert O.lib().orc_pt_decompress_ok(bytes.fromhex(h)) == 0
        assert P.ristretto_decode(bytes.fromhex(h)) is None
    for h in KAT["rfc9496_basepoint_multiples"]:
        assert O.lib().orc_pt_decompress_ok(bytes.fromhex(h)) == 1


def test_rfc9496_one_way_map():
    k = KAT["rfc9496_one_way_map"]
    h = hashlib.sha512(k["label_sha512"].encode()).digest()
    out = C.create_string_buffer(32)
    O.lib().orc_from_uniform_bytes(h, out)
    assert out.raw.hex() == k["encoding"]
    assert P.ristretto_encode(P.from_uniform_bytes(h)).hex() == k["encoding"]


def _sodium():
    for pat in ("/opt/prime-rl/.venv/lib/python3*/site-packages/pyzmq.libs/libsodium*", "/usr/lib/x86_64-linux-gnu/libsodium.so*"):
        for path in glob.glob(pat):
            try:
                s = C.CDLL(path)
                s.sodium_init()
                return s
            except OSError:
                pass
    return None


def test_against_libsodium():
    S = _sodium()
    if S is None:
        pytest.skip("libsodium not present in this image")
    rng = random.Random(3)
    out = C.create_string_buffer(32)
    ref = C.create_string_buffer(32)
    B = bytes.fromhex(KAT["rfc9496_basepoint_multiples"][1])
    for _ in range(20):
        u = bytes(rng.randrange(256) for _ in range(64))
        O.lib().orc_from_uniform_bytes(u, out)
        S.crypto_core_ristretto255_from_hash(ref, u)
        assert out.raw == ref.raw
        k = rng.randrange(L)
        S.crypto_scalarmult_ristretto255_base(ref, k.to_bytes(32, "little"))
        assert O.msm([k], B) == ref.raw
    # a 40-term MSM against libsodium scalarmult + add
    pts, ks = [], []
    for _ in range(40):
        u = bytes(rng.randrange(256) for _ in range(64))
        S.crypto_core_ristretto255_from_hash(out, u)
        pts.append(out.raw)
        ks.append(rng.randrange(L) if rng.random() < 0.8 else rng.randrange(1 << 20))
    acc = None
    for k, p in zip(ks, pts):
        S.crypto_scalarmult_ristretto255(ref, k.to_bytes(32, "little"), p)
        term = ref.raw
        if acc is None:
            acc = term
        else:
            S.crypto_core_ristretto255_add(out, acc, term)
            acc = out.raw
    assert O.msm(ks, b"".join(pts)) == acc


# ------------------------------------------------------------------------------------------------ merlin / keccak / shake (merlin boundary)
def test_keccak_and_merlin_vector():
    st = P.keccak_f1600(bytes(200))
    assert st[:8][::-1].hex() == KAT["keccak_f1600_zero_state_lane0"]
    m = KAT["merlin_test_vector"]
    t = P.Transcript(m["protocol"].encode())
    t.append_message(m["label"].encode(), m["data"].encode())
    assert t.challenge_bytes(m["challenge_label"].encode(), 32).hex() == m["challenge_32"]
    lib = O.lib()
    h = C.c_void_p(lib.orc_transcript_new(m["protocol"].encode(), C.c_uint64(len(m["protocol"]))))
    lib.orc_transcript_append(h, m["label"].encode(), m["data"].encode(), C.c_uint64(len(m["data"])))
    out = C.create_string_buffer(32)
    lib.orc_transcript_challenge(h, m["challenge_label"].encode(), out, C.c_uint64(32))
    lib.orc_transcript_free(h)
    assert out.raw.hex() == m["challenge_32"]


def test_transcript_cpp_matches_pyref_long_run():
    rng = random.Random(4)
    lib = O.lib()
    t = P.Transcript(b"snark_example")
    h = C.c_void_p(lib.orc_transcript_new(b"snark_example", C.c_uint64(13)))
    out = C.create_string_buffer(64)
    for i in range(200):
        msg = bytes(rng.randrange(256) for _ in range(rng.choice([0, 1, 32, 165, 166, 167, 400])))
        label = rng.choice([b"poly_commitment_share", b"C", b"challenge_nextround", b"a"])
        t.append_message(label, msg)
        lib.orc_transcript_append(h, label, msg, C.c_uint64(len(msg)))
        if i % 3 == 0:
            n = rng.choice([32, 64])
            lib.orc_transcript_challenge(h, b"c", out, C.c_uint64(n))
            assert out.raw[:n] == t.challenge_bytes(b"c", n)
    lib.orc_transcript_challenge_scalar(h, b"challenge_tau", out)
    assert int.from_bytes(out.raw[:32], "little") == P.fl_from_bytes_wide(t.challenge_bytes(b"challenge_tau", 64))
    lib.orc_transcript_free(h)


def test_shake256_and_generator_derivation():
    out = C.create_string_buffer(300)
    O.lib().orc_shake256(b"gens_r1cs_sat", C.c_uint64(13), out, C.c_uint64(300))
    assert out.raw == hashlib.shake_256(b"gens_r1cs_sat").digest(300)
    g = KAT["survey_provisional_generators"]
    got = O.derive_gens(g["label"].encode(), 3)
    assert [got[32 * i:32 * i + 32].hex() for i in range(3)] == g["G"]
    # C++ == tier-0 python for two labels, and every stream is a prefix of the longer one
    for label in (b"gens_r1cs_sat", b"gens_r1cs_eval"):
        G, h = P.derive_gens(label, 5)
        cc = O.derive_gens(label, 5)
        assert b"".join(P.ristretto_encode(p) for p in G) + P.ristretto_encode(h) == cc
        assert O.derive_gens(label, 9)[: 32 * 5] == cc[: 32 * 5]
