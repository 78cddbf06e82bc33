"""Multi-rank host logic on CPU: CPI sharding and the detection-list gather over gloo (world_size 2).
The data path itself never communicates; this is the one exchange step."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import rsp_b200 as rsp
from rsp_b200 import stream


def test_shard_range_partitions_exactly():
    for n, world in ((1024, 8), (1024, 3), (5, 8), (0, 2), (17, 4)):
        blocks = [stream.shard_range(n, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        sizes = [hi - lo for lo, hi in blocks]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        stream.shard_range(10, 2, 2)


def _fake_detections(seed, n):
    rng = np.random.default_rng(seed)
    d = np.zeros(n, dtype=rsp.DETECTION_DTYPE)
    d["pair_idx"], d["r_idx"], d["v_idx"] = rng.integers(1, 8, n), rng.integers(16, 5000, n), rng.integers(16, 48, n)
    d["power"], d["range"], d["velocity"], d["angle"] = rng.random(n), rng.random(n) * 1e4, rng.random(n), rng.random(n)
    return d


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = stream.shard_range(7, rank, world)                 # 7 CPIs over 2 ranks -> 4 + 3
    lists = [_fake_detections(100 + i, 10 + i) for i in range(lo, hi)]
    counts, recs = stream.pack_detections(lists, slots=4, cap=32)
    gathered = stream.gather_detections(counts, recs)
    ok = True
    for r in range(world):
        rlo, rhi = stream.shard_range(7, r, world)
        for s, i in enumerate(range(rlo, rhi)):
            want = rsp.sort_detections(_fake_detections(100 + i, 10 + i))
            ok &= bool(np.array_equal(gathered[r][s], want))
        for s in range(rhi - rlo, 4):
            ok &= len(gathered[r][s]) == 0
    q.put((rank, ok))
    dist.destroy_process_group()


def test_gather_detections_world_size_2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]


def test_pack_overflow_is_an_error():
    with pytest.raises(OverflowError):
        stream.pack_detections([_fake_detections(1, 40)], slots=1, cap=32)


def _async_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    slots, cap, maxdet = 8, 16, 32
    ok = True
    # the device ring as CPU tensors: counts [slots] int32 as bytes, records [slots, maxdet * 40] bytes
    lists = [[_fake_detections(1000 * rank + 10 * b + i, 3 + i + b) for i in range(4)] for b in range(2)]     # two batches of 4 CPIs
    counts = torch.zeros(slots, dtype=torch.int32)
    recs = torch.zeros((slots, maxdet * stream.REC_BYTES), dtype=torch.uint8)
    for b in range(2):
        c, r = stream.pack_detections(lists[b], slots=4, cap=maxdet)
        counts[4 * b:4 * b + 4] = c
        recs[4 * b:4 * b + 4] = r
    gat = stream.AsyncDetectionGather(counts.view(torch.uint8), recs, batch_slots=4, cap=cap)
    h0 = gat.launch(0)
    h1 = gat.launch(4)
    out0, out1 = gat.wait(h0), gat.wait(h1)
    for b, out in enumerate((out0, out1)):
        for r in range(world):
            for i in range(4):
                want = rsp.sort_detections(_fake_detections(1000 * r + 10 * b + i, 3 + i + b))
                ok &= bool(np.array_equal(out[r][i], want))
    # more detections than the message holds: an error on every rank, never a truncated list
    counts[0] = cap + 1
    h = gat.launch(0)
    try:
        gat.wait(h)
        ok = False
    except OverflowError:
        pass
    q.put((rank, ok))
    dist.destroy_process_group()


def test_async_packed_gather_world_size_2_gloo():
    """stream.AsyncDetectionGather (one packed all_gather per batch: count header + records) over gloo on CPU tensors."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_async_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]
