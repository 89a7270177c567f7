"""CPU, world_size 2 over gloo: the N>1 plumbing bench.py uses (sharding, barrier, max over ranks, host gather)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    import torch.distributed as dist
    from face_alignment_cvpr_2012_b200 import FACE_DTYPE, sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 101
    lo, hi = sharding.shard_range(n, rank, world)
    local = np.zeros(hi - lo, FACE_DTYPE)
    local["dominant"] = np.arange(lo, hi)
    dist.barrier()
    t = sharding.max_over_ranks(1.0 + rank, dist)
    allr = sharding.gather_records(local, dist)
    iob = np.repeat(np.arange(7), 3)
    mine = sharding.shard_frames(iob, 7, rank, world)
    q.put((rank, lo, hi, t, None if allr is None else allr["dominant"].tolist(), mine.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, t0, all0, m0), (r1, lo1, hi1, t1, all1, m1) = out
    assert (lo0, hi0, lo1, hi1) == (0, 51, 51, 101)
    assert t0 == t1 == 2.0
    assert all0 == list(range(101)) and all1 is None
    assert sorted(m0 + m1) == list(range(21)) and not set(m0) & set(m1)


def test_shard_range_edges():
    from face_alignment_cvpr_2012_b200 import sharding
    for n in (0, 1, 7, 8, 4096):
        for w in (1, 2, 4, 8):
            spans = [sharding.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
