"""Multi-GPU local energy: one process per GPU, the unique-state batch sharded across ranks.

The reference has no distributed code at all (single process); this module adds the only exchange steps the
path needs, with torch.distributed as plumbing (NCCL over NVLink on GPUs, gloo in the CPU tests):

  1. all-gather of the per-rank (key, psi) shards  -> every rank holds the full amplitude lookup table
  2. fused E_loc kernel on the local shard (rows are independent given the table; the Pauli table is replicated)
  3. all-reduce (sum) of five fp64 scalars [sum w, sum w Re E, sum w Im E, sum w (Re E)^2, n]
     (the quantities of src/optimizer/energy.py:328,372-375)

E_loc itself stays sharded: the loss only needs the global mean.
"""
import numpy as np
import torch
import torch.distributed as dist

from .energy import stats_from_sums


def shard_bounds(n, world_size, rank):
    """Contiguous block partition of n rows: rank r owns [lo, hi); the first n % world_size ranks get one extra row."""
    base, extra = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _world(group):
    if not dist.is_available() or not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def gather_table(keys, psi, group=None, equal_sizes=False, out=None):
    """All-gather the (key, psi) shards of every rank -> (keys [T, W] int64 bit patterns, psi [T] complex, T).

    equal_sizes=True skips the size exchange (and its host synchronisation) when every rank is known to hold the same
    number of rows; out=(g_keys, g_psi) reuses preallocated gather buffers.

    Shards may have different lengths: each is padded to the longest for the collective and the padding is REMOVED again
    (by the exchanged counts) before anything is returned — a padded entry must never reach a lookup build: with
    duplicates_equal / assume_unique the build keeps ONE copy of a key with a plain store, so a padded (key, 0) pair could
    replace the real amplitude of that key."""
    world, rank = _world(group)
    keys = keys if keys.dim() == 2 else keys.reshape(-1, 1)
    if world == 1:
        return keys, psi, keys.shape[0]
    if equal_sizes:
        counts = [keys.shape[0]] * world
    else:
        n = torch.tensor([keys.shape[0]], dtype=torch.int64, device=keys.device)
        counts = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(counts, n, group=group)
        counts = [int(c.item()) for c in counts]
    n_max = max(counts)
    if keys.shape[0] < n_max:
        pad = n_max - keys.shape[0]
        keys = torch.cat([keys, torch.zeros((pad, keys.shape[1]), dtype=keys.dtype, device=keys.device)], 0)
        psi = torch.cat([psi, torch.zeros(pad, dtype=psi.dtype, device=psi.device)], 0)
    if out is not None:
        g_keys, g_psi = out
    else:
        g_keys = torch.empty((world * n_max, keys.shape[1]), dtype=keys.dtype, device=keys.device)
        g_psi = torch.empty(world * n_max, dtype=psi.dtype, device=psi.device)
    if keys.is_cuda:
        dist.all_gather_into_tensor(g_keys, keys.contiguous(), group=group)
        dist.all_gather_into_tensor(torch.view_as_real(g_psi), torch.view_as_real(psi.contiguous()), group=group)
    else:  # gloo: list form
        dist.all_gather(list(g_keys.chunk(world, 0)), keys.contiguous(), group=group)
        dist.all_gather(list(torch.view_as_real(g_psi).chunk(world, 0)), torch.view_as_real(psi.contiguous()), group=group)
    total = sum(counts)
    if total != world * n_max:  # uneven shards: drop the padding rows (block r keeps its first counts[r] rows)
        keep = (torch.arange(n_max, device=keys.device)[None, :] < torch.tensor(counts, device=keys.device)[:, None]).reshape(-1)
        g_keys, g_psi = g_keys[keep], g_psi[keep]
    return g_keys, g_psi, total


INT32_MIN = -(2 ** 31)  # bit pattern of -0.0f: "absent" for the MAX all-reduce and a numeric zero for the kernel


def can_allreduce_table(table, psi):
    """The dense complex64 table can itself be all-reduced when the key space is small (N <= 22) and psi is complex64."""
    return table.n_qubits <= 22 and table.words == 1 and psi.dtype == torch.complex64


def aligned_dense_table(table):
    """[2^N, 2] int32 device buffer whose address is a multiple of its size (naqs_lookup_attach_dense32 requires it: the
    kernel forms entry addresses with XORs).  A view into a buffer of twice the size; keep it and pass it back as `out`."""
    n_bytes = (1 << table.n_qubits) * 8
    raw = torch.empty(2 * n_bytes, dtype=torch.uint8, device=table.device)
    off = (-raw.data_ptr()) % n_bytes
    return raw[off:off + n_bytes].view(torch.int32).view(-1, 2)


def allreduce_dense_table(table, keys, psi, group=None, out=None):
    """Multi-GPU lookup build for small key spaces WITHOUT gathering the pairs: every rank scatters its own (key, psi) into a
    direct-address complex64 table pre-filled with -0.0f, the tables are all-reduced with MAX on their int32 bit patterns
    (NCCL over NVLink / NVSwitch, in-switch reduction where available), and the result is attached as the lookup table.
    Volume: 8 * 2^N bytes per rank whatever the number of ranks (an all-gather moves 16 * M * world).  Requires psi to be a
    function of the state (copies of a key on several ranks are identical): the same contract as duplicates_equal.
    out: a buffer from aligned_dense_table(table) to reuse between calls."""
    from . import _lib
    n_entries = 1 << table.n_qubits
    if out is None:
        out = aligned_dense_table(table)
    out.fill_(INT32_MIN)
    k = keys if keys.dim() == 2 else keys.reshape(-1, 1)
    with torch.cuda.device(table.device):
        _lib.check(_lib.load().naqs_dense32_scatter(_lib.ptr(out), n_entries, _lib.ptr(k), _lib.ptr(torch.view_as_real(psi.contiguous())), k.shape[0],
                                                    _lib.stream_ptr(table.device)), "naqs_dense32_scatter")
    world, _ = _world(group)
    if world > 1:
        dist.all_reduce(out, op=dist.ReduceOp.MAX, group=group)
    table.attach_dense32(out)
    return out


def reduce_stats(sums5, group=None):
    """All-reduce (sum) the five statistics sums; returns the reduced tensor (in place)."""
    world, _ = _world(group)
    if world > 1:
        dist.all_reduce(sums5, op=dist.ReduceOp.SUM, group=group)
    return sums5


class Comm:
    """The C-ABI communicator (naqs_comm_t, include/naqs_eloc.h "Multi-GPU exchange"): the table exchange and the statistics
    all-reduce as push kernels over peer memory (key spaces <= 2^22, complex64 psi) or NCCL all-gather + build.
    torch.distributed only carries the 128-byte NCCL unique id from rank 0 to the other ranks."""

    def __init__(self, device, group=None):
        import ctypes as C
        from . import _lib
        self.world, self.rank = _world(group)
        self.device = torch.device(device)
        self.group = group
        lib = _lib.load()
        uid = np.zeros(128, np.uint8)
        if self.rank == 0:
            _lib.check(lib.naqs_comm_unique_id(_lib.ptr(uid)), "naqs_comm_unique_id")
        if self.world > 1:
            backend = dist.get_backend(group)
            t = torch.from_numpy(uid).to(self.device if backend == "nccl" else "cpu")
            dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            uid = t.cpu().numpy()
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(lib.naqs_comm_init(C.byref(self._h), _lib.ptr(np.ascontiguousarray(uid)), self.world, self.rank, self.device.index), "naqs_comm_init")

    def exchange(self, table, keys, psi, max_local=None, flags=0):
        """(key, psi) shards of all ranks -> lookup table of `table` (naqs_table_exchange).  max_local: the largest shard size of
        all ranks — every rank must pass the SAME value (it sizes the gather slots and decides merge vs push, a collective
        decision).  None: it is obtained with one all-reduce (MAX) of the shard sizes; pass it to keep that off the step."""
        from . import _lib
        k = _lib.keys_to_device(keys, table.words, table.device)
        p, code = _lib.psi_to_device(psi, table.device)
        n = k.shape[0]
        if max_local is None and self.world > 1:
            backend = dist.get_backend(self.group)
            m = torch.tensor([n], dtype=torch.int64, device=table.device if backend == "nccl" else "cpu")
            dist.all_reduce(m, op=dist.ReduceOp.MAX, group=self.group)
            max_local = int(m.item())
        with torch.cuda.device(table.device):
            _lib.check(_lib.load().naqs_table_exchange(table._h, self._h, _lib.ptr(k), _lib.ptr(p), code, n, n if max_local is None else int(max_local), flags,
                                                       _lib.stream_ptr(table.device)), "naqs_table_exchange")
        table._lookup_built = True
        return k, p

    def allreduce_stats(self, sums5):
        from . import _lib
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().naqs_stats_allreduce(self._h, _lib.ptr(sums5), _lib.stream_ptr(self.device)), "naqs_stats_allreduce")
        return sums5

    def close(self):
        from . import _lib
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.load().naqs_comm_destroy(self._h)
            self._h.value = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def sharded_local_energy_comm(table, comm, keys_shard, psi_shard, weights_shard=None, out=None, max_local=None, flags=0):
    """One sharded step through the C-ABI exchange: naqs_table_exchange -> naqs_eloc on the shard -> naqs_eloc_stats ->
    naqs_stats_allreduce.  Copies of a key on several ranks must carry the same amplitude (psi is a function of the state)."""
    k, p = comm.exchange(table, keys_shard, psi_shard, max_local=max_local, flags=flags)
    eloc = table.local_energy(k, torch.view_as_complex(p), out=out, rebuild_lookup=False)
    return eloc, comm.allreduce_stats(table.stats(eloc, weights_shard))


def sharded_local_energy(table, keys_shard, psi_shard, weights_shard=None, group=None, out=None, equal_sizes=False, gather_out=None,
                         duplicates_equal=False):
    """E_loc of this rank's shard against the batch of ALL ranks, plus the globally reduced statistics.

    table: DeviceTermTable (replicated on every rank); keys_shard / psi_shard: CUDA tensors (or host arrays) of the
    local unique states and their amplitudes.  -> (eloc_shard float64 [m, 2] on the device, the five reduced sums as a
    device tensor; `stats_from_sums` turns them into mean / variance)."""
    from . import _lib
    k = _lib.keys_to_device(keys_shard, table.words, table.device)
    p, _ = _lib.psi_to_device(psi_shard, table.device)
    pc = torch.view_as_complex(p)
    if duplicates_equal and can_allreduce_table(table, pc):
        allreduce_dense_table(table, k, pc, group)
        eloc = table.local_energy(k, pc, out=out, rebuild_lookup=False)
        return eloc, reduce_stats(table.stats(eloc, weights_shard), group)
    g_keys, g_psi, _ = gather_table(k, pc, group, equal_sizes=equal_sizes, out=gather_out)
    # several ranks may have sampled the same configuration: by default copies are summed (the reference's semantics for a
    # repeated index); duplicates_equal=True states that psi is a function of the state, so one copy is kept
    table.build_lookup(g_keys, g_psi, duplicates_equal=duplicates_equal)
    eloc = table.local_energy(k, pc, out=out, rebuild_lookup=False)
    sums = reduce_stats(table.stats(eloc, weights_shard), group)
    return eloc, sums


def sharded_local_energy_stats(table, keys_shard, psi_shard, weights_shard=None, group=None):
    """Convenience wrapper: -> (eloc_shard, dict(sum_w, mean, variance, n)) with the statistics brought to the host."""
    eloc, sums = sharded_local_energy(table, keys_shard, psi_shard, weights_shard, group)
    return eloc, stats_from_sums(sums.cpu().numpy())
