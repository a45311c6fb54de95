"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path — unique-id exchange, the row partition
(vpin_shard_rows) and the all-gather layout — with the oracle standing in for the per-rank MSM kernels."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import helpers as H
    import oracle_lib as O
    from vpin_b200 import api

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. the id made on rank 0 reaches every rank unchanged
        uid = api.exchange_unique_id(dist, rank, make_id=lambda: bytes(range(128)))
        assert uid == bytes(range(128))
        # 2. every rank commits to its own rows of a 64 x 64 Hyrax grid; one all-gather rebuilds the commitment
        ell = 12
        Lr, R = 1 << (ell // 2), 1 << (ell - ell // 2)
        Z = H.rand_scalars(1 << ell, seed=ell)
        blinds = H.rand_scalars(Lr, seed=50, edge=False)
        r0, r1, sharded = api.shard_rows(Lr, rank, world)
        assert sharded and (r1 - r0) * world == Lr and r0 == rank * (Lr // world)
        gens = O.derive_gens(b"gens_r1cs_sat", R + 1)  # G[0..R) | gens_1 | h
        G, h = gens[: 32 * R], gens[32 * (R + 1): 32 * (R + 2)]
        mine = b"".join(O.msm(Z[i * R:(i + 1) * R] + [blinds[i]], G + h) for i in range(r0, r1))
        buf = torch.zeros(32 * Lr, dtype=torch.uint8)
        buf[32 * r0:32 * r1] = torch.frombuffer(bytearray(mine), dtype=torch.uint8)
        parts = [torch.zeros(32 * (r1 - r0), dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(parts, buf[32 * r0:32 * r1].clone())
        full = b"".join(bytes(p.numpy()) for p in parts)
        assert full == O.hyrax_commit(Z, b"gens_r1cs_sat", blinds, threads=1)
        # 3. the transcript is replayed identically on every rank: a challenge derived from the gathered commitment agrees
        import hashlib
        digest = torch.frombuffer(bytearray(hashlib.sha256(full).digest()), dtype=torch.uint8)
        other = [torch.zeros(32, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(other, digest)
        assert all(bytes(o.numpy()) == bytes(digest.numpy()) for o in other)
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


def test_row_partition_is_exact():
    sys.path.insert(0, ROOT)
    from vpin_b200 import api
    for world in (1, 2, 4, 8):
        for rows in (1, 2, 16, 64, 128, 256, 1024, 4096, 16384):
            spans = [api.shard_rows(rows, r, world) for r in range(world)]
            if all(s[2] for s in spans):
                assert world > 1 and rows // world >= 32
                assert [s[0] for s in spans] == [r * (rows // world) for r in range(world)]
                assert spans[-1][1] == rows and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            else:  # too small (or world 1): every rank keeps the whole range and no collective runs
                assert not any(s[2] for s in spans) and all(s[:2] == (0, rows) for s in spans)


@pytest.mark.timeout(300)
def test_sharded_commit_over_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
