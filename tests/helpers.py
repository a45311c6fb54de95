import random

import numpy as np

import oracle_lib as O

L = O.L_ORDER
EDGE = [0, 1, 2, L - 1, L - 2, (L - 1) // 2, (L + 1) // 2, O.R_MONT, (O.R_MONT * O.R_MONT) % L, 2**128 - 1, 2**252]


def rand_scalars(n, seed, edge=True, small=False):
    rng = random.Random(seed)
    out = [rng.randrange(1 << 20) if small else rng.randrange(L) for _ in range(n)]
    if edge:
        for i, e in enumerate(EDGE):
            if i < n:
                out[(i * 7919) % n] = e
    return out


def synthetic_r1cs(num_cons, num_vars, num_inputs, seed):
    """Satisfiable instance in the style of R1CSInstance::produce_synthetic_r1cs (Spartan/src/r1csinstance.rs:160-238):
    one entry per row in each of A, B, C, with a seeded random assignment (the reference draws it from OsRng)."""
    rng = random.Random(seed)
    size_z = num_vars + num_inputs + 1
    Z = [rng.randrange(L) for _ in range(size_z)]
    Z[num_vars] = 1
    A = np.zeros(num_cons, O.COO_DTYPE)
    B = np.zeros(num_cons, O.COO_DTYPE)
    Cm = np.zeros(num_cons, O.COO_DTYPE)
    one = np.frombuffer(O.le32(1), dtype=np.uint8)
    for i in range(num_cons):
        a_idx, b_idx, c_idx = i % size_z, (i + 2) % size_z, (i + 3) % size_z
        A[i] = (i, a_idx, one)
        B[i] = (i, b_idx, one)
        ab = Z[a_idx] * Z[b_idx] % L
        if Z[c_idx] == 0:
            Cm[i] = (i, num_vars, np.frombuffer(O.le32(ab), dtype=np.uint8))
        else:
            Cm[i] = (i, c_idx, np.frombuffer(O.le32(ab * pow(Z[c_idx], -1, L)), dtype=np.uint8))
    vars_ = O.ints_to_bytes(Z[:num_vars])
    inputs = O.ints_to_bytes(Z[num_vars + 1:])
    # split the assignment like vPIN does: "para" gets every 7th variable, "input" the rest
    para = [z if i % 7 == 0 else 0 for i, z in enumerate(Z[:num_vars])]
    inp = [0 if i % 7 == 0 else z for i, z in enumerate(Z[:num_vars])]
    return A, B, Cm, O.ints_to_bytes(para), O.ints_to_bytes(inp), vars_, inputs
