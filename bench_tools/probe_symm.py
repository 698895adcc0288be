"""Probe (multi-GPU box): NCCL all-reduce latency at the dense-table size and availability of symmetric memory / multicast."""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)


def timeit(fn, n=50):
    for _ in range(10):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


x = torch.zeros((1 << 20, 2), dtype=torch.int32, device=dev)
small = torch.zeros(5, dtype=torch.float64, device=dev)
res = {"allreduce_max_8MB_us": timeit(lambda: dist.all_reduce(x, op=dist.ReduceOp.MAX)),
       "allreduce_sum_40B_us": timeit(lambda: dist.all_reduce(small)),
       "fill_8MB_us": timeit(lambda: x.fill_(-2147483648))}
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty((1 << 20, 2), dtype=torch.int32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    res["symm_mem"] = {"multicast_ptr": int(hdl.multicast_ptr), "world": hdl.world_size, "rank": hdl.rank,
                       "buffer_ptrs": len(hdl.buffer_ptrs), "signal_pad_ptrs": len(hdl.signal_pad_ptrs)}
    res["symm_barrier_us"] = timeit(lambda: hdl.barrier(channel=0))
    f = symm_mem.empty(1 << 21, dtype=torch.float32, device=dev)
    symm_mem.rendezvous(f, dist.group.WORLD.group_name)
    try:
        res["multimem_all_reduce_sum_8MB_us"] = timeit(lambda: torch.ops.symm_mem.multimem_all_reduce_(f, "sum", dist.group.WORLD.group_name))
    except Exception as e:  # noqa: BLE001
        res["multimem_all_reduce_err"] = repr(e)[:300]
    try:
        res["two_shot_all_reduce_sum_8MB_us"] = timeit(lambda: torch.ops.symm_mem.two_shot_all_reduce_(f, "sum", dist.group.WORLD.group_name))
    except Exception as e:  # noqa: BLE001
        res["two_shot_err"] = repr(e)[:300]
except Exception as e:  # noqa: BLE001
    res["symm_mem_err"] = repr(e)[:500]
if rank == 0:
    import json
    print(json.dumps(res))
dist.destroy_process_group()
