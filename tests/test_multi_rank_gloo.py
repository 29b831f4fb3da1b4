"""World-size-2 gloo test (CPU) of the multi-GPU plumbing: instance sharding + the trivial result gather.
Each rank evaluates the host-side palettes of ITS instance range only; the gathered records must reproduce what a
single process computes for the whole crowd (no data-path collective is involved)."""
import os
import socket
import zlib

import numpy as np
import pytest

from reze_engine_b200 import sharding


def test_instance_ranges_partition_the_crowd():
    for K in (1, 7, 8, 4096, 65536, 65537):
        for W in (1, 2, 3, 4, 8):
            spans = [sharding.instance_range(K, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == K
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
            for k in {0, K - 1, K // 2, K // 3}:
                r = sharding.owner_of(k, K, W)
                assert spans[r][0] <= k < spans[r][1]
    with pytest.raises(ValueError):
        sharding.instance_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _crowd_checksums(first, last, K):
    from reze_engine_b200 import synth
    wl = synth.make_workload(64, 24, seed=11)
    world = synth.make_palettes(wl.bones, K, np.random.default_rng(5))     # phase k*0.618 per instance: rank-independent
    return [float(zlib.crc32(world[k].tobytes())) for k in range(first, last)]


def _worker(rank, world_size, port, K, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    first, last = sharding.instance_range(K, world_size, rank)
    sums = _crowd_checksums(first, last, K)
    rec = [float(rank), float(first), float(last), float(sum(sums) % 2**31), 1.5 + rank]
    allrec = sharding.gather_records(rec)
    mx = sharding.max_over_ranks(1.5 + rank)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, allrec, mx, sums))


def test_two_rank_gloo_gather_matches_single_process():
    import torch.multiprocessing as mp
    K, W = 9, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, W, port, K, q)) for r in range(W)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(W))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _crowd_checksums(0, K, K)
    joined = []
    for rank, allrec, mx, sums in got:
        assert mx == 2.5                                   # max over ranks, as bench.py reports time
        assert [r[0] for r in allrec] == [0.0, 1.0]        # every rank sees every record, in rank order
        assert allrec[0][2] == allrec[1][1]                # contiguous ranges
        joined += sums
    assert joined == ref                                   # the shards together are exactly the single-process crowd
    assert got[0][1] == got[1][1]
