"""CPU: the carry-free radix-2^29 F_p multiplier of the MSM hot loop (vpin_b200/csrc/fp29.cuh, host build) against Python big
integers: random operands, the extreme operands of every bound the header states (all limbs at their maximum, both signs),
output form (non-negative limbs within the documented bounds), and the 8 x u32 <-> 9 x 29 conversions."""
import ctypes as C
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 2**255 - 19
N = 1 << 29


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("fp29") / "libfp29.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "cpp", "fp29_shim.cpp")])
    return C.CDLL(so)


def val(limbs):
    return sum(int(v) << (29 * k) for k, v in enumerate(limbs))


def mul(shim, a, b, unsigned):
    out = (C.c_int32 * 9)()
    shim.fp29_mul((C.c_int32 * 9)(*a), (C.c_int32 * 9)(*b), int(unsigned), out)
    return list(out)


def check(shim, a, b, unsigned):
    r = mul(shim, a, b, unsigned)
    assert all(0 <= v < N for v in r[1:]), r
    assert 0 <= r[0] < N + (1 << 24), r
    assert val(r) % P == (val(a) * val(b)) % P


def test_random_and_extreme_products(shim):
    rng = random.Random(29)
    slack = (1 << 24) - 1
    # normal values: limb 0 may carry the slack
    def normal():
        return [rng.randrange(N + slack)] + [rng.randrange(N) for _ in range(8)]
    for _ in range(2000):
        x, y, z, w = normal(), normal(), normal(), normal()
        diff = [p - q for p, q in zip(x, y)]
        summ = [p + q for p, q in zip(z, w)]
        check(shim, diff, z, False)           # (Y - X) * ym
        check(shim, summ, x, False)           # (Y + X) * yp
        check(shim, diff, [p - q for p, q in zip(z, w)], False)   # E * F
        check(shim, diff, summ, False)        # E * H, F * G
        check(shim, summ, [p + q for p, q in zip(x, y)], True)    # G * H
    top = [N + slack - 1] + [N - 1] * 8
    for sa in (1, -1):
        for sb in (1, -1):
            check(shim, [sa * v for v in top], [sb * 2 * v for v in top], False)
            check(shim, [sa * v for v in top], [sb * v for v in top], False)
    # alternating signs (largest cancellation) and sparse operands
    alt = [(-1) ** k * v for k, v in enumerate(top)]
    check(shim, alt, [2 * v for v in alt], False)
    check(shim, alt, [-2 * v for v in alt], False)
    check(shim, [2 * v for v in top], [2 * v for v in top], True)
    check(shim, [0] * 9, top, False)
    check(shim, [0] * 9, top, True)
    check(shim, [1] + [0] * 8, [1] + [0] * 8, True)
    for k in range(9):
        e = [0] * 9
        e[k] = -(N - 1)
        check(shim, e, [2 * v for v in top], False)
        e[k] = 2 * (N - 1)
        check(shim, e, [2 * v for v in top], True)


def test_conversions(shim):
    rng = random.Random(30)
    for _ in range(2000):
        x = rng.randrange(1 << 256) if rng.random() < 0.9 else (1 << 256) - 1 - rng.randrange(64)
        w = [(x >> (32 * j)) & 0xFFFFFFFF for j in range(8)]
        out = (C.c_int32 * 9)()
        shim.fp29_unpack((C.c_uint32 * 8)(*w), out)
        limbs = list(out)
        assert val(limbs) == x and all(0 <= v < N for v in limbs)
        # back: any operand-form value with non-negative limbs packs to something congruent and below 2^256
        big = [rng.randrange(2 * N + (1 << 25)) for _ in range(9)]
        back = (C.c_uint32 * 8)()
        shim.fp29_to_fp((C.c_int32 * 9)(*big), back)
        got = sum(int(v) << (32 * j) for j, v in enumerate(back))
        assert got % P == val(big) % P and got < (1 << 256)
    full = [2 * N + (1 << 25) - 1] * 9
    back = (C.c_uint32 * 8)()
    shim.fp29_to_fp((C.c_int32 * 9)(*full), back)
    assert sum(int(v) << (32 * j) for j, v in enumerate(back)) % P == val(full) % P
