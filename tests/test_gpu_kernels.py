"""Parity of every kernel-level C-ABI entry point with the CPU oracle (bit-exact: integer work)."""
import pytest

import helpers as H
import oracle_lib as O

pytestmark = pytest.mark.gpu


def test_derive_gens_matches_oracle(ctx):
    for label, n in ((b"gens_r1cs_sat", 37), (b"gens_r1cs_eval", 130)):
        assert ctx.derive_gens(label, n) == O.derive_gens(label, n)


@pytest.mark.parametrize("n", [1, 2, 5, 64, 65, 300])
def test_msm_matches_oracle(ctx, n):
    s = H.rand_scalars(n, seed=100 + n)
    gens = O.derive_gens(b"gens_r1cs_eval", n)[: 32 * n]
    assert ctx.msm(b"gens_r1cs_eval", O.ints_to_bytes(s)) == O.msm(s, gens)


def test_msm_small_and_negative_scalars(ctx):
    n = 200
    s = H.rand_scalars(n, seed=7, edge=False, small=True)
    s[3] = O.L_ORDER - 1
    s[4] = O.L_ORDER - 5
    s[9] = 0
    gens = O.derive_gens(b"gens_r1cs_eval", n)[: 32 * n]
    assert ctx.msm(b"gens_r1cs_eval", O.ints_to_bytes(s)) == O.msm(s, gens)


@pytest.mark.parametrize("ell,blinds", [(2, False), (5, True), (8, True), (10, False), (13, True)])
def test_hyrax_commit_matches_oracle(ctx, ell, blinds):
    n = 1 << ell
    Z = H.rand_scalars(n, seed=ell)
    bl = H.rand_scalars(1 << (ell // 2), seed=50 + ell, edge=False) if blinds else None
    got = ctx.hyrax_commit(b"gens_r1cs_sat", O.ints_to_bytes(Z), O.ints_to_bytes(bl) if bl else None)
    assert got == O.hyrax_commit(Z, b"gens_r1cs_sat", bl)


def test_hyrax_commit_is_linear(ctx):
    """size-independent property: commit(Z1 + Z2; b1 + b2) = commit(Z1; b1) + commit(Z2; b2) row-wise
    (the homomorphism vPIN's driver asserts, proof_point_add.rs:69-73)"""
    ell = 12
    n = 1 << ell
    Z1, Z2 = H.rand_scalars(n, 1), H.rand_scalars(n, 2)
    b1, b2 = H.rand_scalars(64, 3, edge=False), H.rand_scalars(64, 4, edge=False)
    Zs = [(a + b) % O.L_ORDER for a, b in zip(Z1, Z2)]
    bs = [(a + b) % O.L_ORDER for a, b in zip(b1, b2)]
    c1 = ctx.hyrax_commit(b"gens_r1cs_sat", O.ints_to_bytes(Z1), O.ints_to_bytes(b1))
    c2 = ctx.hyrax_commit(b"gens_r1cs_sat", O.ints_to_bytes(Z2), O.ints_to_bytes(b2))
    cs = ctx.hyrax_commit(b"gens_r1cs_sat", O.ints_to_bytes(Zs), O.ints_to_bytes(bs))
    assert ctx.commitments_add(c1, c2) == cs


@pytest.mark.parametrize("ell", [0, 1, 3, 9, 12, 13, 17])
def test_eq_evals(ctx, ell):
    r = H.rand_scalars(ell, seed=ell, edge=False)
    got = O.bytes_to_ints(ctx.eq_evals(O.ints_to_bytes(r)))
    assert got == O.eq_evals(r)


@pytest.mark.parametrize("len_", [2, 8, 1024, 1 << 15])
def test_sumcheck_rounds_and_bind(ctx, len_):
    A, B, Cc, D = (H.rand_scalars(len_, seed=s + len_) for s in (1, 2, 3, 4))
    b = lambda v: O.ints_to_bytes(v)
    assert O.bytes_to_ints(ctx.sumcheck_cubic_round(b(A), b(B), b(Cc), b(D))) == O.cubic_round(A, B, Cc, D)
    assert O.bytes_to_ints(ctx.sumcheck_quad_round(b(A), b(B))) == O.quad_round(A, B)
    assert O.bytes_to_ints(ctx.sumcheck_cubic3_round(b(A), b(B), b(Cc))) == O.cubic3_round(A, B, Cc)
    r = H.rand_scalars(1, seed=99, edge=False)[0]
    assert O.bytes_to_ints(ctx.bind_top(b(A), O.le32(r))) == O.bind_top(A, r)


def test_sumcheck_claim_consistency(ctx):
    """size-independent property at a larger size: after binding with r, eval_0 + eval_1 of the next round equals
    the previous round polynomial at r (the verifier's check, sumcheck.rs:46)."""
    n = 1 << 16
    A, B = H.rand_scalars(n, 11), H.rand_scalars(n, 12)
    L = O.L_ORDER
    e0, e2 = O.bytes_to_ints(ctx.sumcheck_quad_round(O.ints_to_bytes(A), O.ints_to_bytes(B)))
    claim = sum(a * b for a, b in zip(A, B)) % L
    e1 = (claim - e0) % L
    # quadratic through (0,e0),(1,e1),(2,e2)
    r = 123456789
    inv2 = pow(2, -1, L)
    a = inv2 * (e2 - 2 * e1 + e0) % L
    bq = (e1 - e0 - a) % L
    at_r = (a * r * r + bq * r + e0) % L
    A2 = O.bytes_to_ints(ctx.bind_top(O.ints_to_bytes(A), O.le32(r)))
    B2 = O.bytes_to_ints(ctx.bind_top(O.ints_to_bytes(B), O.le32(r)))
    assert sum(x * y for x, y in zip(A2, B2)) % L == at_r


@pytest.mark.parametrize("ell", [4, 7, 12])
def test_bound(ctx, ell):
    n = 1 << ell
    Z = H.rand_scalars(n, seed=ell)
    Lv = H.rand_scalars(1 << (ell // 2), seed=ell + 1, edge=False)
    assert O.bytes_to_ints(ctx.bound(O.ints_to_bytes(Z), O.ints_to_bytes(Lv))) == O.bound(Z, Lv)


def test_spmv_and_transpose(ctx):
    from vpin_b200 import api

    num_cons, num_vars, num_inputs = 256, 128, 3
    A, B, Cm, _, _, vars_, inputs = H.synthetic_r1cs(num_cons, num_vars, num_inputs, seed=5)
    # add a heavy column (constant-1 column referenced by every row) and duplicate entries
    import numpy as np

    extra = np.zeros(num_cons + 2, O.COO_DTYPE)
    for i in range(num_cons):
        extra[i] = (i, num_vars, np.frombuffer(O.le32(i + 5), dtype=np.uint8))
    extra[num_cons] = (3, 7, np.frombuffer(O.le32(O.L_ORDER - 1), dtype=np.uint8))
    extra[num_cons + 1] = (3, 7, np.frombuffer(O.le32(9), dtype=np.uint8))
    # every code of the value dictionary (+-1, +-2, 3: no multiplication, value never read) next to general values in short rows,
    # and two long rows (> kLongRow entries: the warp-per-row kernel): a 128-term bit decomposition 2^i like vPIN's, and a
    # mixed one of 40 terms
    L = O.L_ORDER
    dict_vals = [1, L - 1, 2, L - 2, 3, 4, L - 3, 0]
    more = [(10 + k, (5 * k + 1) % num_vars, v) for k, v in enumerate(dict_vals)] + [(10 + k, (3 * k + 2) % num_vars, dict_vals[(k + 3) % 8]) for k in range(8)]
    more += [(40, i % (num_vars + 1 + num_inputs), 1 << i) for i in range(128)]
    more += [(41, (7 * i) % num_vars, dict_vals[i % 8] if i % 3 else 12345 + i) for i in range(40)]
    extra2 = np.zeros(len(more), O.COO_DTYPE)
    for k, (r, c, v) in enumerate(more):
        extra2[k] = (r, c, np.frombuffer(O.le32(v), dtype=np.uint8))
    A2 = np.concatenate([A, extra, extra2])
    inst = api.Instance(ctx, num_cons, num_vars, num_inputs, A2, B, Cm)
    z = O.bytes_to_ints(vars_) + [1] + O.bytes_to_ints(inputs)
    z += [0] * (2 * num_vars - len(z))
    got = inst.spmv_abc(O.ints_to_bytes(z))
    for g, M in zip(got, (A2, B, Cm)):
        assert O.bytes_to_ints(g) == O.spmv(M, num_cons, 2 * num_vars, z)
    x = H.rand_scalars(num_cons, seed=77)
    got = inst.spmv_t_abc(O.ints_to_bytes(x))
    for g, M in zip(got, (A2, B, Cm)):
        assert O.bytes_to_ints(g) == O.spmv(M, num_cons, 2 * num_vars, x, transposed=True)


def test_instance_errors(ctx):
    """error behaviour of Instance::new (Spartan/src/lib.rs:649-692 tests)"""
    import numpy as np
    from vpin_b200 import api

    one = np.frombuffer(O.le32(1), dtype=np.uint8)
    ok = np.array([(0, 0, one)], dtype=O.COO_DTYPE)
    bad_idx = np.array([(100, 0, one)], dtype=O.COO_DTYPE)
    with pytest.raises(api.VpinError) as e:
        api.Instance(ctx, 4, 4, 1, bad_idx, ok, ok)
    assert e.value.name == "InvalidIndex"
    big = np.frombuffer(bytes([0xFF] * 32), dtype=np.uint8)
    bad_val = np.array([(0, 0, big)], dtype=O.COO_DTYPE)
    with pytest.raises(api.VpinError) as e:
        api.Instance(ctx, 4, 4, 1, ok, bad_val, ok)
    assert e.value.name == "InvalidScalar"


# ---- fused sumcheck rounds (kernels_round.cu) against the plain definition, big-int arithmetic --------------------------
def _sumcheck_reference(degree, tables, r):
    """SP/sumcheck.rs:619-676 / :456-486 with the challenges given: per round eval at 0, 2(, 3), then bound_poly_var_top"""
    p = O.L_ORDER
    T = [list(t) for t in tables]
    evals = []
    for rj in r:
        half = len(T[0]) // 2
        pts = (0, 2, 3) if degree == 3 else (0, 2)
        for t in pts:
            acc = 0
            for i in range(half):
                v = [(tab[i] + t * (tab[half + i] - tab[i])) % p for tab in T]
                acc += v[0] * (v[1] * v[2] - v[3]) if degree == 3 else v[0] * v[1]
            evals.append(acc % p)
        T = [[(tab[i] + rj * (tab[half + i] - tab[i])) % p for i in range(half)] for tab in T]
    return evals, [tab[0] for tab in T]


@pytest.mark.parametrize("degree,ell", [(3, 1), (3, 2), (3, 7), (3, 11), (2, 1), (2, 5), (2, 12)])
def test_fused_sumcheck_rounds(ctx, degree, ell):
    n = 1 << ell
    ntab = 4 if degree == 3 else 2
    tables = [H.rand_scalars(n, seed=1000 * degree + 10 * ell + k) for k in range(ntab)]
    r = H.rand_scalars(ell, seed=77 + ell, edge=False)
    evals, finals = ctx.sumcheck_fused(degree, [O.ints_to_bytes(t) for t in tables], O.ints_to_bytes(r))
    want_evals, want_finals = _sumcheck_reference(degree, tables, r)
    assert O.bytes_to_ints(evals) == want_evals
    assert O.bytes_to_ints(finals) == want_finals


# ---- SPARK timestamps (kernels_sort.cu) against the reference's sequential replay ------------------------------------
def _replay(addrs, N, M):
    """AddrTimestamps::new, SP/sparse_mlpoly.rs:237-257: one counter per cell, shared by the three matrices"""
    audit = [0] * M
    addr_out, ts_out = [], []
    for a in addrs:
        for i in range(N):
            x = int(a[i]) if i < len(a) else 0  # padding entries are (0, 0, 0) and do bump address 0 (:370-378)
            addr_out.append(x)
            ts_out.append(audit[x])
            audit[x] += 1
    return addr_out, ts_out, audit


@pytest.mark.parametrize("N,M,sizes,hot", [(8, 4, (8, 5, 0), 0.0), (4096, 512, (4096, 4000, 1), 0.5), (8192, 1 << 17, (8000, 8192, 7777), 0.2),
                                           (1 << 15, 1 << 9, (1 << 15, 30000, 12345), 0.9), (16, 1 << 20, (0, 0, 0), 0.0)])
def test_spark_timestamps_match_the_sequential_replay(ctx, N, M, sizes, hot):
    import numpy as np
    rng = np.random.default_rng(N + M)
    addrs = []
    for n in sizes:
        a = rng.integers(0, M, size=n, dtype=np.uint32)
        a[rng.random(n) < hot] = min(3, M - 1)  # a hot cell, like the constant-1 column of vPIN's instances
        addrs.append(a)
    got_addr, got_ts, got_audit = ctx.spark_timestamps(addrs, N, M)
    want_addr, want_ts, want_audit = _replay(addrs, N, M)
    assert got_addr.tolist() == want_addr
    assert got_ts.tolist() == want_ts
    assert got_audit.tolist() == want_audit


@pytest.mark.parametrize("k", [1, 2, 5, 11])
def test_product_tree_layers(ctx, k):
    """ProductCircuit::new (Spartan/src/product_tree.rs:18-56): every layer is the element-wise product of the two halves of the
    layer below, down to the two factors of the root (big-int restatement; edge values 0, 1, l - 1 among the leaves)."""
    n = 1 << k
    leaves = H.rand_scalars(n, seed=40 + k)
    want, layer = list(leaves), list(leaves)
    while len(layer) > 2:
        half = len(layer) // 2
        layer = [layer[i] * layer[i + half] % H.L for i in range(half)]
        want += layer
    got = O.bytes_to_ints(ctx.product_tree(O.ints_to_bytes(leaves)))
    assert len(got) == 2 * n - 2 and got == want


@pytest.mark.parametrize("n", [1, 33, 4096])
def test_hash_layer_and_deref_gather(ctx, n):
    """deref_mem (Spartan/src/sparse_mlpoly.rs:267-276) and build_hash_layer (:547-622): h = ts gamma^2 + val gamma + addr - tau,
    the write set one timestamp later; addresses and timestamps at their extremes (0, 2^32 - 1)."""
    import random
    rng = random.Random(n)
    cells = 64
    mem = H.rand_scalars(cells, seed=n + 1)
    addr = [rng.randrange(cells) for _ in range(n)]
    vals = O.bytes_to_ints(ctx.deref_gather(addr, O.ints_to_bytes(mem)))
    assert vals == [mem[a] for a in addr]
    big_addr = [rng.choice([0, 1, 2**32 - 1, rng.randrange(2**32)]) for _ in range(n)]
    ts = [rng.choice([0, 1, 2**32 - 1, rng.randrange(1 << 20)]) for _ in range(n)]
    gamma, tau = rng.randrange(H.L), rng.randrange(H.L)
    rd, wr = ctx.hash_layer(big_addr, O.ints_to_bytes(vals), ts, O.le32(gamma), O.le32(tau))
    g2 = gamma * gamma % H.L
    want = [(t * g2 + v * gamma + a - tau) % H.L for a, v, t in zip(big_addr, vals, ts)]
    assert O.bytes_to_ints(rd) == want
    assert O.bytes_to_ints(wr) == [(w + g2) % H.L for w in want]
