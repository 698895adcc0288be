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
    for rank, keys, psi, g_keys, g_psi, total, sums, red in res:
        # no padding survives the gather (a padded (key, 0) pair could replace a real amplitude in a keep-one lookup build)
        assert total == sum(sizes) and g_keys.shape == (total, 1) and g_psi.shape == (total,)
        off = 0
        for r, (_, k_r, p_r, *_rest) in enumerate(res):
            assert np.array_equal(g_keys[off:off + sizes[r]], k_r) and np.array_equal(g_psi[off:off + sizes[r]], p_r)
            off += sizes[r]
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


def _maxtrick_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_keys = 64
        rng = np.random.default_rng(7)                       # same generator on every rank: psi is a function of the key
        psi_all = (rng.normal(size=n_keys) + 1j * rng.normal(size=n_keys)).astype(np.complex64)
        psi_all[3] = np.complex64(complex(-0.0, 1.5))        # a component that IS -0.0 (collides with the "absent" pattern)
        mine = np.random.default_rng(50 + rank).permutation(n_keys)[:40]      # overlapping shards
        tbl = torch.full((n_keys, 2), nd.INT32_MIN, dtype=torch.int32)
        bits = torch.from_numpy(np.ascontiguousarray(psi_all[mine]).view(np.float32).reshape(-1, 2).view(np.int32).copy())
        tbl[torch.from_numpy(mine)] = bits                   # what naqs_dense32_scatter does on the device
        dist.all_reduce(tbl, op=dist.ReduceOp.MAX)
        q.put((rank, mine, tbl.numpy().copy(), psi_all))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_allreduce_max_on_float_bit_patterns_world2():
    """The dense-table all-reduce: MAX over int32 bit patterns with -0.0f (INT32_MIN) as "absent" reproduces the union of
    the ranks' (key, psi) sets exactly — including negative amplitudes — and absent entries stay numeric zeros."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_maxtrick_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    union = np.union1d(res[0][1], res[1][1])
    psi_all = res[0][3]
    for _, _, tbl, _ in res:
        got = tbl.view(np.float32).view(np.complex64).reshape(-1)
        assert np.array_equal(got[union], psi_all[union])                    # value-equal (-0.0 == 0.0)
        absent = np.setdiff1d(np.arange(64), union)
        assert np.all(got[absent] == 0)
        assert np.array_equal(res[0][2], tbl)                                 # every rank ends with the same table
