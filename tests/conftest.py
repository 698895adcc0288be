import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_table(mol):
    d = np.load(os.path.join(GOLDEN, "tables", f"{mol}.npz"))
    N, na, nb = (int(x) for x in d["meta"])
    return d["xy"], d["yz"], d["coeff"], N, na, nb


def load_case(name):
    return dict(np.load(os.path.join(GOLDEN, f"eloc_{name}.npz")))


def load_terms_json(mol):
    with open(os.path.join(GOLDEN, f"terms_{mol}.json")) as f:
        raw = json.load(f)
    return {tuple((q, p) for q, p in t): complex(re, im) for t, re, im in raw}


CASE_TABLE = {"LiH_sector": "LiH", "LiH_small": "LiH", "H2O_sector": "H2O", "NH3_1000": "NH3", "N2_2000": "N2",
              "N2_1.5_500": "N2_1.5", "N2_full_3000": "N2", "LiH_full_600": "LiH",
              # 30 qubits, the largest molecule of the paper: produced by the reference's own _HilbertRestricted (2^30-entry LUT, paid once)
              "Li2O_500": "Li2O"}
# the same seeded batches through PauliHamiltonian.get(dtype=np.float32), the reference constructor's default
F32_CASE_TABLE = {"LiH_sector_f32": "LiH", "H2O_300_f32": "H2O", "N2_full_1000_f32": "N2"}


def case_sector(case, mol):
    """(n_alpha, n_beta) the fixture was generated with (None, None for the full space)."""
    meta = case["meta"]
    return (None, None) if int(meta[1]) < 0 else (int(meta[1]), int(meta[2]))


def random_sector_states(N, na, nb, m, seed):
    """m distinct random keys with na bits on even and nb bits on odd qubits."""
    rng = np.random.default_rng(seed)
    ev, od = np.arange(0, N, 2), np.arange(1, N, 2)
    seen, out = set(), []
    while len(out) < m:
        v = sum(1 << int(q) for q in rng.choice(ev, na, replace=False)) | sum(1 << int(q) for q in rng.choice(od, nb, replace=False))
        if v not in seen:
            seen.add(v)
            out.append(v)
    return np.array(out, dtype=np.uint64)


@pytest.fixture(scope="session")
def built_library():
    import importlib
    b = importlib.import_module("naqs-for-quantum-chemistry_b200._build")
    return b.build_library()
