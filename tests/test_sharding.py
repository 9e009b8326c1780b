"""Frame-pair sharding: partition arithmetic, and the one collective of the multi-GPU path -- the
all-gather of per-frame object records -- exercised with two gloo ranks on the CPU."""
import os
import socket

import numpy as np
import pytest

from moving_object_detector_b200 import OBJECT_DTYPE, STATS_DTYPE
from moving_object_detector_b200.sharding import (merge_gathered, padded_block, shard_frames, shard_range)


@pytest.mark.parametrize("n_pairs,world", [(1024, 8), (63, 2), (63, 4), (7, 8), (0, 2), (1, 1), (1000, 3)])
def test_shards_are_a_contiguous_partition(n_pairs, world):
    cover = []
    sizes = []
    for r in range(world):
        a, b = shard_range(n_pairs, r, world)
        assert 0 <= a <= b <= n_pairs
        cover += list(range(a, b))
        sizes.append(b - a)
        fa, fb = shard_frames(n_pairs, r, world)
        assert (fb - fa) == (b - a + 1 if b > a else 0)
    assert cover == list(range(n_pairs))
    assert max(sizes) - min(sizes) <= 1 and max(sizes) == padded_block(n_pairs, world)


def test_bad_shard_requests_raise():
    for args in ((10, 2, 2), (10, -1, 2), (10, 0, 0), (-1, 0, 1)):
        with pytest.raises(ValueError):
            shard_range(*args)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


MAX_OBJS = 4


def _rank_records(rank, world, n_pairs):
    """Deterministic fake clusterer output of one rank: pair g has (g % 3) objects whose id/centre encode g."""
    first, last = shard_range(n_pairs, rank, world)
    block = padded_block(n_pairs, world)
    objs = np.zeros((block, MAX_OBJS), OBJECT_DTYPE)
    stats = np.zeros(block, STATS_DTYPE)
    for k, g in enumerate(range(first, last)):
        n = g % 3
        stats["n_objects"][k] = n
        stats["n_clusters"][k] = n
        for j in range(n):
            objs[k, j]["id"] = j
            objs[k, j]["center"] = (g, j, rank)
            objs[k, j]["n_points"] = 100 + g
    return objs, stats


def _worker(rank, world, port, n_pairs, q):
    import torch
    import torch.distributed as dist
    from moving_object_detector_b200.sharding import gather_object_lists
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        objs, stats = _rank_records(rank, world, n_pairs)
        t_objs = torch.from_numpy(objs.view(np.uint8).reshape(objs.shape[0], -1).copy())
        t_stats = torch.from_numpy(stats.view(np.uint8).reshape(stats.shape[0], -1).copy())
        g_objs, g_stats = gather_object_lists(t_objs, t_stats)
        merged = merge_gathered(g_objs.numpy(), g_stats.numpy(), n_pairs, OBJECT_DTYPE, STATS_DTYPE)
        ok = len(merged) == n_pairs
        for g, rec in enumerate(merged):
            ok &= len(rec) == g % 3
            ok &= all(int(r["n_points"]) == 100 + g and float(r["center"][0]) == g for r in rec)
        # the benchmark's form of the same exchange: one all-gather of a preallocated flat buffer, two steps in a row
        from moving_object_detector_b200.sharding import ObjectListGatherer
        gat = ObjectListGatherer(t_objs.shape[0], t_objs.shape[1], t_stats.shape[1], "cpu")
        for step in range(2):
            gat.wait()
            gat.objs.copy_(t_objs)
            gat.stats.copy_(t_stats)
            gat.launch()
        gat.wait()
        o2, s2 = gat.result()
        ok &= bool(torch.equal(o2, g_objs) and torch.equal(s2, g_stats))
        # a ring of three step slots gathered by ONE collective (a stream of steps, one gather per batch); slot k
        # carries the records with every id shifted by k, and a partial ring (two of three slots) gathers a prefix
        ring = ObjectListGatherer(t_objs.shape[0], t_objs.shape[1], t_stats.shape[1], "cpu", ring=3)
        for slots in (3, 2):
            for k in range(slots):
                o, st = ring.slot(k)
                shifted = objs.copy()
                shifted["id"] += k
                o.copy_(torch.from_numpy(shifted.view(np.uint8).reshape(objs.shape[0], -1).copy()))
                st.copy_(t_stats)
            before = ring.launches
            ring.launch(slots)
            ring.wait()
            ok &= ring.launches == before + 1
            for k in range(slots):
                ok_o, ok_s = ring.result(k)
                ok &= bool(torch.equal(ok_s, g_stats))
                for r in range(world):
                    want = _rank_records(r, world, n_pairs)[0]
                    want["id"] += k
                    ok &= bool(np.array_equal(ok_o[r].numpy().tobytes(), want.tobytes()))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [7, 8])
def test_object_list_all_gather_two_gloo_ranks(n_pairs):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0, "a rank failed"
    results = [q.get(timeout=10) for _ in procs]
    assert sorted(results) == [(0, True), (1, True)]
