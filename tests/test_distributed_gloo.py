"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard partition, padded all-gather of the (key, psi)
table, all-reduce of the E_loc statistics.  The kernels themselves need a GPU (tests/test_gpu_parity.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import naqs_b200
from naqs_b200 import distributed as nd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, sizes, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(100 + rank)
        n = sizes[rank]
        keys = torch.from_numpy(rng.integers(1, 2 ** 40, size=(n, 1), dtype=np.int64))
        psi = torch.from_numpy((rng.normal(size=n) + 1j * rng.normal(size=n)).astype(np.complex64))
        g_keys, g_psi, total = nd.gather_table(keys, psi)
        # statistics: local sums -> all-reduce
        e = rng.normal(size=n) + 1j * rng.normal(size=n)
        w = rng.random(n)
        sums = torch.tensor([w.sum(), (w * e.real).sum(), (w * e.imag).sum(), (w * e.real ** 2).sum(), float(n)], dtype=torch.float64)
        red = nd.reduce_stats(sums.clone())
        q.put((rank, keys.numpy(), psi.numpy(), g_keys.numpy(), g_psi.numpy(), total, sums.numpy(), red.numpy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("sizes", [(5, 5), (7, 3), (0, 4)])
def test_gather_table_and_reduce_stats_world2(sizes):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, sizes, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n_max = max(sizes)
    for rank, keys, psi, g_keys, g_psi, total, sums, red in res:
        assert total == sum(sizes) and g_keys.shape == (world * n_max, 1) and g_psi.shape == (world * n_max,)
        for r, (_, k_r, p_r, *_rest) in enumerate(res):
            blk_k, blk_p = g_keys[r * n_max:(r + 1) * n_max], g_psi[r * n_max:(r + 1) * n_max]
            assert np.array_equal(blk_k[:sizes[r]], k_r) and np.array_equal(blk_p[:sizes[r]], p_r)
            assert np.all(blk_p[sizes[r]:] == 0)  # padding carries zero amplitude -> adds nothing to the lookup table
        assert np.allclose(red, res[0][6] + res[1][6], rtol=1e-15)
    st = naqs_b200.stats_from_sums(res[0][7])
    assert st["n"] == sum(sizes)


def test_shard_bounds_partition():
    for n in (0, 1, 7, 100, 1_000_003):
        for world in (1, 2, 3, 8):
            b = [nd.shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_is_identity():
    keys = torch.arange(6, dtype=torch.int64).reshape(-1, 1)
    psi = torch.ones(6, dtype=torch.complex64)
    g_keys, g_psi, total = nd.gather_table(keys, psi)
    assert total == 6 and torch.equal(g_keys, keys) and torch.equal(g_psi, psi)
    s = torch.arange(5, dtype=torch.float64)
    assert torch.equal(nd.reduce_stats(s.clone()), s)
