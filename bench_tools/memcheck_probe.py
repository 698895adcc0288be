"""Probe for the wide-key hash walk with the companion filter: all rows against the C oracle, `reps` times.
   usage: python bench_tools/memcheck_probe.py [reps]   (run plain, or under compute-sanitizer --launch-timeout 0)"""
import sys
import numpy as np
import torch
torch.zeros(1, device="cuda")  # CUDA up before the long host set-up (compute-sanitizer attaches at the first API call)
sys.path.insert(0, ".")
import naqs_b200
from oracle import eloc_oracle as eo
from oracle import c_oracle

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
N, K, M, extra = 70, 400, 150000, 2000
xy, yz, c = eo.synthetic_table(N, K, seed=N + K)
st = eo.synthetic_states(N, M, seed=N + 1)
psi = eo.synthetic_psi(M, seed=K + 1)
ct = c_oracle.COracleTable(xy, yz, c, N)
_, cols, _ = ct.rows(st[:extra])
tk = np.unique(np.concatenate([st, cols]), axis=0)
tp = eo.synthetic_psi(len(tk), seed=7)
ref = ct.local_energy(st, psi, tk, tp)
t = naqs_b200.DeviceTermTable(xy, yz, c, N)
for r in range(reps):
    e = naqs_b200._lib.complex_from_pairs(t.local_energy(st, psi, table_keys=tk, table_psi=tp, kind=naqs_b200._lib.LOOKUP_HASH))
    err = np.abs(e - ref) / np.abs(ref)
    bad = np.nonzero(err > 1e-12)[0]
    print(f"rep {r}: table {len(tk)} bad rows {len(bad)} max err {err.max():.3e} first bad {bad[:12].tolist()}", flush=True)
