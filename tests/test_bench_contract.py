"""bench.py contract checks that need no GPU: the reference arm prints one well-formed JSON line (CPU path of the
reference, compiled kernels from oracle/_ref), and the B200 arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

import pytest
import torch
from conftest import ROOT

REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "e2e", "cpu_baseline", "impl"}


@pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "src", "utils")), reason="oracle/_ref not built")
def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-sample", "3000"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d)
    assert d["impl"] == "reference" and d["metric"] == "eloc_state_term_couplings_per_sec" and d["unit"] == "couplings/s"
    assert d["value"] > 1e6 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == os.cpu_count()
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_b200_arm_fails_loudly_without_gpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3", "--cpu-sample", "0"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_bench_helpers():
    """Workload helpers of bench.py that the extras / --strong legs rely on (no GPU)."""
    import importlib.util
    import numpy as np
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    sec = b.full_sector(12, 2, 2)                       # LiH: C(6,2)^2 = 225 states (hilbert.py:446-449)
    assert len(sec) == 225 and len(np.unique(sec)) == 225
    even = sum(1 << q for q in range(0, 12, 2))
    assert all(bin(int(k) & even).count("1") == 2 and bin(int(k) & (even << 1)).count("1") == 2 for k in sec)
    st = b.sector_states(30, 7, 7, 1000, 3)            # random distinct Li2O sector states
    assert len(st) == 1000 and len(np.unique(st)) == 1000
    ev30 = sum(1 << q for q in range(0, 30, 2))
    assert all(bin(int(k) & ev30).count("1") == 7 and bin(int(k) & (ev30 << 1)).count("1") == 7 for k in st[:50])
    xy, yz, c = b.synthetic_table(70, 120)
    assert xy.shape == (120, 2) and yz.shape == (120, 2) and c.shape == (120,)
    assert not (xy & yz).any()                          # a flipped qubit never carries a phase bit in the generator
    # --strong: key-range shards for direct-address key spaces, contiguous blocks otherwise; the shards tile the batch
    from naqs_b200.distributed import shard_bounds
    full = b.make_workload("n2_1e6", 0, m_override=4000)
    parts = [b.strong_shard(b.make_workload("n2_1e6", 0, m_override=4000), 4, r, shard_bounds) for r in range(4)]
    keys = [np.asarray(p["states"]).reshape(-1) for p in parts]
    assert all(len(k) == 1000 for k in keys) and all(keys[r].max() < keys[r + 1].min() for r in range(3))
    assert np.array_equal(np.sort(np.concatenate(keys)), np.sort(np.asarray(full["states"]).reshape(-1)))
    for p in parts:  # psi travels with its key
        assert np.array_equal(p["psi"], b.psi_of_keys(p["states"], full["N"]))
    li = [b.strong_shard(b.make_workload("li2o_1e5", 0, m_override=600), 2, r, shard_bounds) for r in range(2)]
    assert np.array_equal(np.concatenate([np.asarray(p["states"]).reshape(-1) for p in li]), b.make_workload("li2o_1e5", 0, m_override=600)["states"].reshape(-1))
