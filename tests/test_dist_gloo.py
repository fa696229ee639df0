"""CPU, world_size 2, gloo: the N>1 host logic — contiguous env slabs and the episode-stat all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spacefortress_b200 import dist as sfdist


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, l, w = sfdist.init_from_env(backend="gloo")
    first, count = sfdist.shard(1000003)
    vec = torch.arange(24, dtype=torch.int64) * (rank + 1)
    vec[20] = 10 + rank  # max-reduced slot
    out = sfdist.all_reduce_episode_stats(vec.clone())
    q.put((rank, first, count, out.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_covers_everything_without_overlap():
    for total in (1, 7, 4096, 1000003):
        for world in (1, 2, 3, 8):
            spans = [sfdist.shard(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1


def test_all_reduce_episode_stats_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, f0, c0, o0), (r1, f1, c1, o1) = res
    assert (f0, c0, f1, c1) == (0, 500002, 500002, 500001)
    expect = [k * 3 for k in range(24)]
    expect[20] = 11
    assert o0 == expect and o1 == expect


def test_all_reduce_is_noop_without_process_group():
    v = torch.arange(24, dtype=torch.int64)
    assert torch.equal(sfdist.all_reduce_episode_stats(v.clone()), v)
