"""world_size-2 gloo tests (CPU) of the N>1 path: stream partition, record all-gather, max-over-ranks timing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stella_vslam_b200 import multi_gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = multi_gpu.assign_streams(8, world, rank)
        # per-stream records: (stream id, keypoints, matches)
        rec = torch.tensor([[s, 2000 + s, 1300 + 10 * s] for s in mine], dtype=torch.int32)
        allrec = multi_gpu.gather_records(rec, world)
        tmax = multi_gpu.max_over_ranks(10.0 + rank, torch.device("cpu"), world)
        q.put((rank, mine, allrec.tolist(), tmax))
    finally:
        dist.destroy_process_group()


def test_partition_is_disjoint_and_complete():
    for world in (1, 2, 3, 8):
        seen = sorted(s for r in range(world) for s in multi_gpu.assign_streams(8, world, r))
        assert seen == list(range(8))
        for s in range(8):
            assert s in multi_gpu.assign_streams(8, world, multi_gpu.owner_of(s, world))
    with pytest.raises(ValueError):
        multi_gpu.assign_streams(8, 2, 2)


def test_gather_and_timing_world2():
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
    (r0, m0, a0, t0), (r1, m1, a1, t1) = res
    assert m0 == [0, 2, 4, 6] and m1 == [1, 3, 5, 7]
    assert a0 == a1                                   # every rank sees every stream's record
    assert [row[0] for row in a0[0]] == m0 and [row[0] for row in a0[1]] == m1
    assert a0[1][2] == [5, 2005, 1350]
    assert t0 == t1 == 11.0                           # max over ranks
    assert multi_gpu.frames_per_second(64, 10, 2, 1000.0) == 1280.0


def test_world1_needs_no_process_group():
    rec = torch.arange(6, dtype=torch.int32).reshape(3, 2)
    assert multi_gpu.gather_records(rec, 1).shape == (1, 3, 2)
    assert multi_gpu.max_over_ranks(3.5, torch.device("cpu"), 1) == 3.5
