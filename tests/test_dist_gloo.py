"""N>1 host logic on CPU: world_size 2 over gloo (shard assignment, statistics reduction, max-over-ranks timing)."""
import os
import socket

import numpy as np
import pytest

from projectd_core_b200 import dist as pdist


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 4096, 65536 * 8 + 3):
        for world in (1, 2, 3, 8):
            spans = [pdist.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (o0, c0), (o1, _) in zip(spans, spans[1:]):
                assert o0 + c0 == o1
            if total:
                for g in {0, total // 2, total - 1}:
                    r = pdist.env_rank(g, total, world)
                    o, c = spans[r]
                    assert o <= g < o + c
    with pytest.raises(ValueError):
        pdist.shard_range(8, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, total_envs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    r, local, w = pdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    off, cnt = pdist.shard_range(total_envs, rank, world)
    # per-rank statistics as pd_env_stats would return them: counts derived from the rank's own global env ids
    ids = np.arange(off, off + cnt)
    stats = np.array([cnt, float(ids.sum()), 2.0 * cnt, (ids % 3 == 0).sum(), (ids % 5 == 0).sum(), 0, 0, 0], np.float64)
    red = pdist.reduce_stats(stats)
    ms = pdist.max_over_ranks(10.0 + rank)
    q.put((rank, off, cnt, red.tolist(), ms))
    dist.barrier(); dist.destroy_process_group()


def test_world2_gloo_stats_reduction_and_timing():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue(); port = _free_port(); total = 1001
    ps = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60); assert p.exitcode == 0
    ids = np.arange(total)
    want = [total, float(ids.sum()), 2.0 * total, float((ids % 3 == 0).sum()), float((ids % 5 == 0).sum()), 0, 0, 0]
    for rank, off, cnt, red, ms in out:
        assert red == want            # both ranks hold the global statistics
        assert ms == 11.0             # max over ranks
    assert out[0][1] == 0 and out[0][1] + out[0][2] == out[1][1] and out[1][1] + out[1][2] == total
    s = pdist.summarize(want)
    assert s["episodes"] == total and s["mean_length"] == 2.0
