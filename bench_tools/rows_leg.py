"""Stored-rows leg alone (full N2 sector -> CSR with restricted column indices), for an ncu launch list:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/rows_launches.csv python bench_tools/rows_leg.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import naqs_b200  # noqa: E402

wl = bench.load_table("N2")
xy, yz, c, N, na, nb = wl
t = naqs_b200.DeviceTermTable(xy, yz, c, N, na, nb)
sec = torch.from_numpy(bench.full_sector(N, na, nb).view(np.int64)).cuda()
for _ in range(2):
    out = t.rows(sec)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    out = t.rows(sec)
torch.cuda.synchronize()
print("rows ms", 1e3 * (time.perf_counter() - t0) / 5, "nnz", int(out[0][-1].item()))
