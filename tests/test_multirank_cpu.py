"""world_size-2 `gloo` test of the N>1 host logic (runs on CPU): query sharding, result gather, the
max-over-ranks timing reduction used by bench.py, and output-slab decomposition of the sweeps."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import convitp_b200 as ci
    sh = ci.sharding
    n = 100_003
    rng = np.random.default_rng(0)  # same seed on every rank = replicated inputs
    coords = [rng.random(n) for _ in range(3)]
    fake_eval = lambda x, y, z: torch.from_numpy(x + 2 * y - z)  # stands in for itp.eval (no GPU here)

    def gather(local, lo, hi, total):
        per = sh.shard_bounds(total, world, 0)[1]
        buf = torch.zeros(per, dtype=local.dtype)
        buf[: hi - lo] = local
        out = [torch.zeros(per, dtype=local.dtype) for _ in range(world)]
        dist.all_gather(out, buf)
        return torch.cat(out)[:total]

    full, (lo, hi) = sh.eval_sharded(fake_eval, coords, world, rank, gather)
    ok = bool(np.array_equal(full.numpy(), coords[0] + 2 * coords[1] - coords[2]))
    slow = sh.max_over_ranks_ms(10.0 + 5.0 * rank, world)
    ret[rank] = (ok, lo, hi, slow)
    dist.destroy_process_group()


def test_query_sharding_gather_and_timing_reduce_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret[0][0] and ret[1][0]
    assert ret[0][1] == 0 and ret[0][2] == ret[1][1] and ret[1][2] == 100_003  # disjoint, covering
    assert ret[0][2] % 32 == 0
    assert ret[0][3] == ret[1][3] == 15.0  # max over ranks


@pytest.mark.parametrize("n,world", [(0, 4), (1, 8), (31, 2), (1000, 8), (10 ** 9, 8), (10 ** 9 + 7, 3)])
def test_shard_bounds_cover_and_balance(n, world):
    from convitp_b200 import sharding as sh
    spans = [sh.shard_bounds(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for a, b in zip(spans, spans[1:]):
        assert a[1] == b[0]
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(s for s in sizes if s or True) <= max(sizes)  # trailing ranks may be short / empty
    assert max(sizes) <= -(-n // world) + 31


def test_output_slabs_tile_the_last_axis():
    from convitp_b200 import sharding as sh
    kout = (np.linspace(0, 1, 17), np.linspace(0, 1, 23), np.linspace(0, 1, 101))
    got = []
    for r in range(4):
        ks, (lo, hi) = sh.output_slab(kout, 4, r)
        assert ks[0] is kout[0] and len(ks[2]) == hi - lo
        got.append(ks[2])
    assert np.array_equal(np.concatenate(got), kout[2])
