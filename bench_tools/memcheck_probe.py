"""Memcheck probe for the hash walk with wide keys: run under compute-sanitizer, one variant per argv[1]."""
import sys
import numpy as np
sys.path.insert(0, ".")
import naqs_b200
from oracle import eloc_oracle as eo
from oracle import c_oracle

N, K, M, extra = {"v1": (70, 400, 150000, 2000), "v2": (70, 400, 100000, 0), "v3": (100, 400, 700000, 0), "v4": (40, 600, 150000, 2000),
                  "v5": (70, 400, 30000, 0)}[sys.argv[1]]
xy, yz, c = eo.synthetic_table(N, K, seed=N + K)
st = eo.synthetic_states(N, M, seed=N + 1)
psi = eo.synthetic_psi(M, seed=K + 1)
t = naqs_b200.DeviceTermTable(xy, yz, c, N)
if extra:
    _, cols, _ = c_oracle.COracleTable(xy, yz, c, N).rows(st[:extra])
    tk = np.unique(np.concatenate([st, cols]), axis=0)
    tp = eo.synthetic_psi(len(tk), seed=7)
    e = t.local_energy(st, psi, table_keys=tk, table_psi=tp, kind=naqs_b200._lib.LOOKUP_HASH)
else:
    e = t.local_energy(st, psi, kind=naqs_b200._lib.LOOKUP_HASH)
print(sys.argv[1], N, K, M, "table", len(tk) if extra else M, "sum", float(e.sum().item()))
