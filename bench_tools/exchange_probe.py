"""Time naqs_table_exchange alone, per mode, under torchrun (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench_tools/exchange_probe.py

N2 table (2^20 complex64 entries = 8 MB), 10^6 distinct rows per rank (dense shards) or 10^4 (sparse shards).  Every iteration:
barrier, then CUDA events around the exchange on the rank's stream; reported: median over iterations, max over ranks.  Also the
raw peer-copy bandwidth between GPU 0 and GPU 1 (what the P2P kernels ride on)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import naqs_b200  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
xy, yz, c, N, _, _ = bench.load_table("N2")
table = naqs_b200.DeviceTermTable(xy, yz, c, N, None, None, device=dev)
comm = naqs_b200.distributed.Comm(dev)
out = {"world": world}
for tag, M in (("dense_1e6", 1_000_000), ("sparse_1e4", 10_000)):
    st = np.random.default_rng(rank).choice(2 ** N, M, replace=False).astype(np.uint64)
    d_k = torch.from_numpy(st.view(np.int64)).to(dev).reshape(-1, 1)
    d_p = torch.from_numpy(bench.psi_of_keys(st, N)).to(dev)
    for mode, flags in (("merge", 0x8000), ("nccl_allreduce_max", 0x4000), ("push", 0x2000)):
        ts = []
        for it in range(25):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            comm.exchange(table, d_k, d_p, max_local=M, flags=flags)
            e1.record()
            e1.synchronize()
            if it >= 5:
                ts.append(e0.elapsed_time(e1))
        t = torch.tensor([float(np.median(ts))], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[f"{tag}_{mode}_ms"] = float(t.item())
if rank == 0 and world > 1:
    a = torch.empty(64 << 20, dtype=torch.uint8, device="cuda:0")
    b = torch.empty(64 << 20, dtype=torch.uint8, device="cuda:1")
    for nbytes in (8 << 20, 64 << 20):
        for _ in range(3):
            b[:nbytes].copy_(a[:nbytes])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            b[:nbytes].copy_(a[:nbytes])
        e1.record()
        e1.synchronize()
        out[f"peer_copy_{nbytes >> 20}MB_GBs"] = nbytes * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    out["can_access_peer_0_1"] = bool(torch.cuda.can_device_access_peer(0, 1))
dist.barrier()
if rank == 0:
    print(json.dumps(out))
comm.close()
dist.destroy_process_group()
