#!/usr/bin/env python
"""bench.py — E_loc state·term couplings/sec on B200 (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload n2_1e6|h2o_1e5|li2o_1e5|synthetic]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU path (oracle/_ref Cython kernels) on the host cores

A "step" = one pass of the hot path over one batch: (all-gather of the (key, psi) shards when N > 1) ->
amplitude-lookup build -> fused E_loc kernel -> E_loc statistics (-> all-reduce of 5 fp64 scalars when N > 1).
`value`  : couplings/s (M·K per step, summed over ranks) with the inputs already resident in HBM, CUDA-event timed.
`e2e`    : the same metric through the host-buffer entry (pinned host states + psi in, E_loc out), copies inside the timed region.
Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
METRIC = "eloc_state_term_couplings_per_sec"
UNIT = "couplings/s"


# ----------------------------------------------------------------------------------------- workloads
def load_table(mol):
    d = np.load(os.path.join(GOLDEN, "tables", f"{mol}.npz"))
    N, na, nb = (int(x) for x in d["meta"])
    return d["xy"], d["yz"], d["coeff"], N, na, nb


def psi_for(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.normal(size=n) + 1j * rng.normal(size=n)).astype(np.complex64)


def psi_of_keys(keys, n_qubits, seed=0):
    """psi as a FUNCTION of the state (normal + 1j*normal, complex64, seeded): the same configuration gets the same amplitude
    on every rank, as a wavefunction does.  For small spaces a 2^N table, otherwise a counter-based generator."""
    keys = np.asarray(keys).reshape(-1)
    if n_qubits <= 24:
        rng = np.random.default_rng(seed)
        full = (rng.normal(size=2 ** n_qubits) + 1j * rng.normal(size=2 ** n_qubits)).astype(np.complex64)
        return full[keys.astype(np.int64)]
    x = (keys.astype(np.uint64) + np.uint64(seed)) * np.uint64(0x9E3779B97F4A7C15)
    x ^= x >> np.uint64(29); x *= np.uint64(0xBF58476D1CE4E5B9); x ^= x >> np.uint64(32)
    u1 = ((x >> np.uint64(11)).astype(np.float64) + 0.5) / 2.0 ** 53
    y = x * np.uint64(0x94D049BB133111EB); y ^= y >> np.uint64(31)
    u2 = ((y >> np.uint64(11)).astype(np.float64) + 0.5) / 2.0 ** 53
    r = np.sqrt(-2.0 * np.log(u1))
    return (r * np.cos(2 * np.pi * u2) + 1j * r * np.sin(2 * np.pi * u2)).astype(np.complex64)


def sector_states(N, na, nb, m, seed):
    rng = np.random.default_rng(seed)
    ev, od = np.arange(0, N, 2), np.arange(1, N, 2)
    a = np.zeros(2 * m, np.int64)
    for arr, k in ((ev, na), (od, nb)):
        picks = np.argsort(rng.random((2 * m, len(arr))), axis=1)[:, :k]
        a |= (1 << arr[picks]).sum(axis=1)
    a = np.unique(a)
    return rng.permutation(a)[:m].astype(np.uint64)


def full_sector(N, na, nb):
    """Every key with na bits on even and nb bits on odd qubits (small sectors only), ascending."""
    import itertools
    ev = [sum(1 << q for q in c) for c in itertools.combinations(range(0, N, 2), na)]
    od = [sum(1 << q for q in c) for c in itertools.combinations(range(1, N, 2), nb)]
    return np.sort(np.array([a | b for a in ev for b in od], dtype=np.uint64))


def synthetic_table(N, K, seed=0):
    """Random Pauli sum of SURVEY.md §8d config 5: 2-/4-qubit flip masks shared by ~6 terms each (+ a diagonal group),
    YZ masks = JW-like contiguous run XOR a random weight-N/4 mask, normal fp64 coefficients."""
    rng = np.random.default_rng(seed)
    W = 1 if N <= 63 else 2
    G = max(2, K // 6)
    flips = [0]
    for _ in range(G - 1):
        qs = rng.choice(N, 2 if rng.random() < 0.3 else 4, replace=False)
        flips.append(sum(1 << int(q) for q in qs))
    xy, yz = [], []
    for k in range(K):
        f = flips[k % G]
        lo, hi = sorted(int(q) for q in rng.choice(N, 2, replace=False))
        run = ((1 << hi) - 1) ^ ((1 << lo) - 1)
        rnd = int.from_bytes(np.packbits(rng.random(8 * ((N + 7) // 8)) < 0.25).tobytes(), "little") & ((1 << N) - 1)
        xy.append(f)
        yz.append((run ^ rnd) & ~f)
    m64 = (1 << 64) - 1
    to_words = lambda vals: np.array([[(v >> (64 * w)) & m64 for w in range(W)] for v in vals], dtype=np.uint64)  # noqa: E731
    return to_words(xy), to_words(yz), rng.normal(size=K)


def make_workload(name, rank, m_override=None, synth=None):
    """-> dict(table args, states uint64 [M], psi complex64 [M], description).  Seeded; rank r draws seed r."""
    if name == "n2_1e6":
        xy, yz, c, N, _, _ = load_table("N2")
        M = m_override or 1_000_000
        st = np.random.default_rng(rank).choice(2 ** N, M, replace=False).astype(np.uint64)
        desc = f"N2 STO-3G (20 qubits, K=2239, Kxy=378), M={M} distinct keys of the unrestricted 2^20 space per GPU (seed=rank), sector filter off, psi = seeded complex64 function of the key (SURVEY.md §8d config 3)"
        return dict(xy=xy, yz=yz, c=c, N=N, na=None, nb=None, states=st, psi=psi_of_keys(st, N), desc=desc, mol="N2")
    if name == "h2o_1e5":
        xy, yz, c, N, _, _ = load_table("H2O")
        M = m_override or 100_000
        st = np.random.default_rng(rank).integers(0, 2 ** N, M).astype(np.uint64)
        desc = f"H2O STO-3G (14 qubits, K=1390), M={M} rows drawn with replacement from 2^14 (1e5 unique 14-qubit states do not exist), table = distinct keys (SURVEY.md §8d config 2)"
        return dict(xy=xy, yz=yz, c=c, N=N, na=None, nb=None, states=st, psi=psi_of_keys(st, N), desc=desc, mol="H2O", dedup_table=True)
    if name == "li2o_1e5":
        xy, yz, c, N, na, nb = load_table("Li2O")
        M = m_override or 100_000
        st = sector_states(N, na, nb, M, rank)
        desc = f"Li2O STO-3G (30 qubits, K=20558, Kxy=3810), M={M} distinct (7,7)-sector states per GPU, sector filter on (SURVEY.md §8d config 4 batch)"
        return dict(xy=xy, yz=yz, c=c, N=N, na=na, nb=nb, states=st, psi=psi_of_keys(st, N), desc=desc, mol="Li2O")
    if name == "synthetic":
        N, K, M = synth
        xy, yz, c = synthetic_table(N, K)
        rng = np.random.default_rng(rank)
        W = 1 if N <= 63 else 2
        st = np.zeros((M, W), np.uint64)
        bits = np.argsort(rng.random((M, N)), axis=1)[:, : N // 2]
        for w in range(W):
            sel = (bits >= 64 * w) & (bits < 64 * (w + 1))
            st[:, w] = np.where(sel, np.uint64(1) << np.where(sel, bits - 64 * w, 0).astype(np.uint64), np.uint64(0)).sum(axis=1, dtype=np.uint64)
        st = np.unique(st, axis=0)
        st = st[rng.permutation(len(st))]
        desc = f"synthetic Pauli sum: N={N} qubits, K={K} terms (Kxy~K/6), M={len(st)} random weight-N/2 keys per GPU (SURVEY.md §8d config 5)"
        return dict(xy=xy, yz=yz, c=c, N=N, na=None, nb=None, states=st, psi=psi_for(len(st), rank), desc=desc, mol="synthetic")
    raise SystemExit(f"unknown workload {name}")


def strong_shard(wl, world, rank, shard_bounds):
    """--strong: ONE batch split over the ranks (in place).  Direct-address key spaces (N <= 26) get KEY-RANGE shards — the batch in
    ascending key order, equal row counts per rank — so that the key-order walk skips every 32-key task without a row of this rank
    and its work shrinks with the shard; hash-lookup batches keep contiguous blocks of the generated order."""
    if wl["N"] <= 26:
        order = np.argsort(np.asarray(wl["states"]).reshape(-1), kind="stable")
        wl["states"], wl["psi"] = np.asarray(wl["states"])[order], wl["psi"][order]
        wl["desc"] += " — rows in ascending key order (key-range shards)"
    lo, hi = shard_bounds(len(wl["states"]), world, rank)
    if (hi - lo) * world != len(wl["states"]):
        raise SystemExit("--strong needs a batch size divisible by the number of ranks")
    wl["states"], wl["psi"] = wl["states"][lo:hi], wl["psi"][lo:hi]
    wl["desc"] += f" — STRONG scaling: the batch is split over {world} ranks"
    return wl


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU (NVML) while the timed region runs."""

    def __init__(self, index, period=0.02):
        self.index, self.period, self.samples, self.reasons = index, period, [], set()
        self._stop = threading.Event()
        self._t = None
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(statistics.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def ncu_summary():
    """Per-workload ncu figures of the hot kernel (dram bytes per launch, L1TEX / issue utilisation), profiles/ncu_summary_rNN.json."""
    d = os.path.join(ROOT, "profiles")
    for name in sorted(os.listdir(d), reverse=True) if os.path.isdir(d) else []:
        if name.startswith("ncu_summary") and name.endswith(".json"):
            with open(os.path.join(d, name)) as f:
                return json.load(f)
    return {}


def pipe_peaks():
    """Integer / fp64 pipe ceilings measured with bench_tools/pipe_peaks.cu on this pool's B200 (profiles/)."""
    for name in sorted(os.listdir(os.path.join(ROOT, "profiles")), reverse=True) if os.path.isdir(os.path.join(ROOT, "profiles")) else []:
        if name.startswith("pipe_peaks") and name.endswith(".json"):
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return json.load(f), name
    return None, None


# ----------------------------------------------------------------------------------------- reference arm
# ---------------------------------------------------------------------------------------------- pipe roofline (live)
# Algorithmic work of the two fused formulations per WARP UNIT (32 states x one record of the term stream), DESIGN.md §6:
# (warp instructions, L1TEX wavefronts).  The counts are those of the formulation, not of the compiled loop — loop
# control, address arithmetic beyond one LOP3 / IADD per access, queue traffic of the hash walk's survivors and bank
# conflicts are NOT in them, so frac = t_roof / t_measured charges all of that to the kernel.
#   key-order walk (dense complex64 table, keyorder.cuh): parity word = 2 LDS + 1 XOR; flip masks 2 LDS.128;
#     per group: entry offset 2, LUT LDS.64, zero test, address LOP3, predicated LDG.64, 2 F2F, 2 DFMA = 10 instr,
#     1 (16-entry LUT) or 2 (64-entry LUT) shared wavefronts + 1 table line
#   hash walk, light pass (sliced.cuh): parity word = NN LDS + NN/2 XOR; 6 LDS.128 (hashes, zero masks; B: +1 LDS.64);
#     per group: zero-mask rotate 2, filter offset LOP3, LDS, bit XOR, rotate, 2 LOP3, predicated OR = 9 instr, 1 wavefront;
#     push: 8 instr, 2 wavefronts.  Survivors (data dependent, ~1.5 % of the pairs at 1e5 keys) are not counted.
#   big-group word (30 terms): parity word + header, then per 6-term chunk: offset 2, LUT LDS.64, DADD = 4 instr, 2 wavefronts
#   XU pipe (key-order walk only): the two F2F.F64.F32 of every table read run at 16 lanes per clock and SM = 0.5 warp instructions
ISSUE_PER_CLK_SM, WAVEFRONTS_PER_CLK_SM, XU_PER_CLK_SM = 4.0, 1.0, 0.5  # sm_100a: 4 warp schedulers, one L1TEX data-stage wavefront per clock (ncu peaks)


def stream_units(xy_words):
    """(A records, B records, big-group words, big groups) of a term table, as the stream builders cut it (sliced.cuh / keyorder.cuh)."""
    a = np.ascontiguousarray(xy_words).reshape(len(xy_words), -1)
    _, counts = np.unique(a, axis=0, return_counts=True)
    n_a, n_b, big = int((counts <= 4).sum()), int(((counts > 4) & (counts <= 6)).sum()), counts[counts > 6]
    return (n_a + 7) // 8, (n_b + 4) // 5, int(((big + 29) // 30).sum()), int(len(big))


def pipe_roofline(mode, units, n_warp_units, nn, kernel_ms, sm_mhz, sm_count=148):
    """t_roof = max(issue, L1TEX) time of the algorithmic work above at the SM clock measured DURING the run; frac = t_roof / t_kernel."""
    rec_a, rec_b, words_c, blobs = units
    xu = 0
    if mode == "keyorder":
        per = {"A": (5 + 8 * 10, 4 + 8 * 2), "B": (4 + 5 * 10, 4 + 5 * 3), "C": (4 + 5 * 4, 3 + 5 * 2), "blob": (8, 1)}
        xu = 2 * (8 * rec_a + 5 * rec_b + blobs)
    else:
        par_i, par_w = nn + nn // 2, nn
        per = {"A": (par_i + 6 + 8 * 9 + 8, par_w + 6 + 8 + 2), "B": (par_i + 7 + 5 * 11 + 8, par_w + 7 + 5 + 2),
               "C": (par_i + 5 * 4, par_w + 5 * 2), "blob": (20, 3)}
    instr = rec_a * per["A"][0] + rec_b * per["B"][0] + words_c * per["C"][0] + blobs * per["blob"][0]
    waves = rec_a * per["A"][1] + rec_b * per["B"][1] + words_c * per["C"][1] + blobs * per["blob"][1]
    clk = sm_mhz * 1e6
    t_issue = n_warp_units * instr / (ISSUE_PER_CLK_SM * sm_count * clk)
    t_l1 = n_warp_units * waves / (WAVEFRONTS_PER_CLK_SM * sm_count * clk)
    t_xu = n_warp_units * xu / (XU_PER_CLK_SM * sm_count * clk)
    t_roof = max(t_issue, t_l1, t_xu)
    return {"bound": "xu" if t_xu >= max(t_l1, t_issue) else ("l1tex" if t_l1 >= t_issue else "issue"), "t_roof_ms": 1e3 * t_roof, "t_issue_ms": 1e3 * t_issue,
            "t_l1tex_ms": 1e3 * t_l1, "t_xu_ms": 1e3 * t_xu, "xu_instr_per_unit": int(xu),
            "frac": 1e3 * t_roof / kernel_ms, "warp_units": int(n_warp_units), "warp_instr_per_unit": int(instr), "wavefronts_per_unit": int(waves),
            "formulation": mode, "sm_mhz": sm_mhz, "sm_count": sm_count}


def run_reference(args):
    """The reference's own CPU implementation of the path (compiled Cython kernels from oracle/_ref + the numpy/scipy
    orchestration of hamiltonian.py:272-370 restated in oracle/ref_path.py), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import ref_path
    wl = make_workload(args.workload, 0, m_override=args.ref_sample)
    M, K = len(wl["states"]), len(wl["c"])
    path = ref_path.ReferencePath(wl["xy"], wl["yz"], wl["c"], wl["N"], wl["na"], wl["nb"])
    st = wl["states"].astype(np.int64)
    times = []
    for i in range(args.warmup + args.steps):
        path.reset()  # cold H cache: first-seen states, the steady state for large molecules
        t0 = time.perf_counter()
        path.local_energy(st, wl["psi"])
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = M * K * len(times) / total
    cores = os.cpu_count()
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "sample": f"{M} of the workload's states per step, cold H cache each step"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                             "sample": f"{M} states x K={K} per step; reference Cython kernels (oracle/_ref) + numpy/scipy orchestration, OMP threads={cores}"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def cpu_baseline_sample(wl, sample, reps=2):
    from oracle import ref_path
    M = min(sample, len(wl["states"]))
    path = ref_path.ReferencePath(wl["xy"], wl["yz"], wl["c"], wl["N"], wl["na"], wl["nb"])
    st, psi = wl["states"][:M].astype(np.int64).reshape(-1), wl["psi"][:M]
    best, eloc = None, None
    for _ in range(reps):
        path.reset()
        t0 = time.perf_counter()
        eloc = path.local_energy(st, psi)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    K = len(wl["c"])
    return {"value": M * K / best, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
            "sample": f"first {M} states of the workload, K={K}, cold H cache, best of {reps}: {best:.2f} s "
                      f"(reference Cython kernels from oracle/_ref + numpy/scipy orchestration of hamiltonian.py:272-370)"}, (st, psi, eloc)


def measure_extras(naqs_b200, dev, args):
    """The other BASELINE.json configs on one GPU (device-resident couplings/s, same timing rules, fewer steps), and the
    E_loc call of a LiH VMC iteration (config 0) through the reference-shaped host API next to the reference CPU path."""
    import torch
    out = {}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for name in ("h2o_1e5", "li2o_1e5"):
        if name == args.workload:
            continue
        wl = make_workload(name, 0)
        table = naqs_b200.DeviceTermTable(wl["xy"], wl["yz"], wl["c"], wl["N"], wl["na"], wl["nb"], device=dev)
        M, K = len(wl["states"]), table.K
        d_states = torch.from_numpy(np.ascontiguousarray(wl["states"]).reshape(M, -1).view(np.int64)).to(dev)
        d_psi = torch.from_numpy(wl["psi"]).to(dev)
        d_out = torch.empty((M, 2), dtype=torch.float64, device=dev)
        if wl.get("dedup_table"):
            uk, first = np.unique(wl["states"], return_index=True)
            tk, tp = torch.from_numpy(uk.view(np.int64)).to(dev).reshape(-1, 1), d_psi[torch.from_numpy(first).to(dev)]
        else:
            tk, tp = d_states, d_psi
        ms = []
        for i in range(5 + 20):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            table.build_lookup(tk, tp, assume_unique=True)
            table.local_energy(d_states, d_psi, out=d_out, rebuild_lookup=False)
            table.stats(d_out)
            e1.record()
            e1.synchronize()
            if i >= 5:
                ms.append(e0.elapsed_time(e1))
        out[name] = {"value": M * K * len(ms) / (sum(ms) * 1e-3), "unit": UNIT, "ms_per_step": sum(ms) / len(ms), "steps": len(ms),
                     "workload": wl["desc"]}
        del table
    # config 0: the E_loc call inside a LiH VMC iteration (<= 225 unique states): latency of one call
    xy, yz, c, N, na, nb = load_table("LiH")
    case = np.load(os.path.join(GOLDEN, "eloc_LiH_sector.npz"))
    st, psi = case["states"].astype(np.int64).astype(np.int16), case["psi"]
    table = naqs_b200.DeviceTermTable(xy, yz, c, N, na, nb, device=dev)
    for _ in range(5):
        table.local_energy_host(st, psi, assume_unique=True, out_dtype=np.complex64)
    t0 = time.perf_counter()
    n = 200
    for _ in range(n):
        table.local_energy_host(st, psi, assume_unique=True, out_dtype=np.complex64)
    b200_ms = 1e3 * (time.perf_counter() - t0) / n
    lih = {"states": int(len(st)), "terms": int(len(c)), "b200_host_call_ms": b200_ms}
    try:
        from oracle import ref_path
        path = ref_path.ReferencePath(xy, yz, c, N, na, nb)
        cold = []
        for _ in range(5):
            path.reset()
            t0 = time.perf_counter()
            path.local_energy(st.astype(np.int64), psi)
            cold.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        for _ in range(20):
            path.local_energy(st.astype(np.int64), psi)  # warm H cache: only get_H + sparse_dense_mv
        warm = (time.perf_counter() - t0) / 20
        lih.update({"reference_cpu_cold_cache_ms": 1e3 * min(cold), "reference_cpu_warm_cache_ms": 1e3 * warm, "cores": os.cpu_count()})
    except Exception as e:  # noqa: BLE001
        lih["reference_cpu"] = f"unavailable: {e}"
    out["lih_vmc_eloc_call"] = lih
    # CSR-rows mode (the update_H replacement: stored couplings with restricted column indices, bit-exact matrix elements):
    # the full N2 sector (14 400 states -> 1.3 M stored elements), device result in device memory vs the reference's update_H
    try:
        xy, yz, c, N, na, nb = load_table("N2")
        sec = full_sector(N, na, nb)
        table = naqs_b200.DeviceTermTable(xy, yz, c, N, na, nb, device=dev)
        d_sec = torch.from_numpy(sec.view(np.int64)).to(dev).reshape(-1, 1)
        for _ in range(20):  # the GPU idled through the CPU reference leg above: let the clocks come back up
            indptr, _, _, _ = table.rows(d_sec)
        rows_ms = float("inf")
        for _ in range(3):   # best of three batches of ten calls (a 0.2 ms host-synchronous call is sensitive to host noise)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                indptr, cols, ridx, vals = table.rows(d_sec)
            torch.cuda.synchronize()
            rows_ms = min(rows_ms, 1e3 * (time.perf_counter() - t0) / 10)
        rows = {"states": int(len(sec)), "nnz": int(indptr[-1].item()), "b200_rows_ms": rows_ms}
        try:
            from oracle import ref_path
            path = ref_path.ReferencePath(xy, yz, c, N, na, nb)
            best = 1e30
            for _ in range(2):
                path.reset()
                t0 = time.perf_counter()
                H = path.update_H(sec.astype(np.int64), check_unseen=False, assume_unique=True)
                best = min(best, time.perf_counter() - t0)
            rows.update({"reference_cpu_update_H_ms": 1e3 * best, "reference_nnz": int(H.nnz), "cores": os.cpu_count()})
        except Exception as e:  # noqa: BLE001
            rows["reference_cpu"] = f"unavailable: {e}"
        out["n2_sector_csr_rows"] = rows
    except Exception as e:  # noqa: BLE001
        out["n2_sector_csr_rows"] = {"error": repr(e)}
    # configs 1 / 4: a VMC iteration of the reference's own loop (experiments/_base._run, flags of batch_train.sh:14) per backend
    if not args.no_vmc:
        out["vmc_iteration"] = vmc_iteration_split()
    return out


def vmc_iteration_split(molecules=(("LiH", 12, 100000, 10000), ("N2", 12, 1000000, 10000)), timeout=900):
    """ms per VMC iteration, split sample / state2idx / E_loc / rest (backward + optimizer step), for
       reference   : the reference's compiled Cython E_loc on the host cores (its own loop, unmodified)
       b200        : naqs_b200.install() — same loop, E_loc through the C ABI with host buffers
       b200_device : install(device_resident=True, fused_loss=True) — sampler output handed over on the GPU, fused loss terms
    The model runs on the GPU in all three (the reference's default).  oracle/ref_vmc.py drives the loop in a subprocess per
    backend from the staged reference tree (oracle/_ref/tree); the first 2 iterations are dropped as warm-up."""
    import subprocess
    import tempfile
    from oracle import ref_vmc
    if ref_vmc.tree_root() is None:
        return {"unavailable": "no reference tree (oracle/_ref/tree is staged by __graft_entry__.build())"}
    res = {}
    for mol, iters, n_samps, n_unq_min in molecules:
        entry = {"iterations": iters - 2, "n_samps": n_samps, "cores": os.cpu_count()}
        for label, backend, extra in (("reference", "reference", ()), ("b200", "b200", ()), ("b200_device", "b200", ("--device-resident", "--fused-loss"))):
            with tempfile.TemporaryDirectory() as td:
                outp = os.path.join(td, "r.npz")
                cmd = [sys.executable, "-m", "oracle.ref_vmc", "--molecule", mol, "--iters", str(iters), "--seed", "111", "--backend", backend,
                       "--out", outp, "--no-solve", "--n-samps", str(n_samps), "--n-unq-min", str(n_unq_min), *extra]
                env = dict(os.environ, NAQS_ELOC_BACKEND=backend)
                try:
                    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
                    if r.returncode != 0:
                        entry[label] = {"error": r.stderr[-400:]}
                        continue
                    rec = ref_vmc.load_record(outp)
                except Exception as e:  # noqa: BLE001
                    entry[label] = {"error": repr(e)}
                    continue
            ms = {k: 1e3 * float(np.mean(rec[k][2:])) if len(rec[k]) > 2 else None for k in ("t_sample", "t_state2idx", "t_eloc", "t_step")}
            total = sum(v for k, v in ms.items() if v is not None and k != "t_eloc")
            entry[label] = {"sample_ms": ms["t_sample"], "state2idx_ms": ms["t_state2idx"], "eloc_ms": ms["t_eloc"],
                            "backward_step_ms": (ms["t_step"] - ms["t_eloc"]) if ms["t_step"] is not None and ms["t_eloc"] is not None else None,
                            "iteration_ms": total, "unique_states": int(np.mean([len(i) for i in rec["idx"][2:]])) if len(rec["idx"]) > 2 else None,
                            "last_energy": rec["log_eloc"][-1] if rec["log_eloc"] else None}
        res[mol] = entry
    return res


# ----------------------------------------------------------------------------------------- B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("BENCH_WL", "n2_1e6"))
    ap.add_argument("--states", type=int, default=None, help="override M (states per GPU)")
    ap.add_argument("--synthetic", type=int, nargs=3, default=[64, 10000, 100000], metavar=("N", "K", "M"))
    ap.add_argument("--cpu-sample", type=int, default=100_000, help="states of the bounded cpu_baseline sample (0 = skip)")
    ap.add_argument("--ref-sample", type=int, default=50_000, help="states per step of --impl reference")
    ap.add_argument("--strong", action="store_true", help="strong scaling: ONE batch of M states (seed 0) split across the ranks (default: weak, M per rank)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configs and the LiH call-latency leg")
    ap.add_argument("--no-vmc", action="store_true", help="skip the VMC-iteration split of the extras (three subprocess runs per molecule)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import naqs_b200
    from naqs_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch N>1 with torch.distributed.run (one process per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    wl = make_workload(args.workload, 0 if args.strong else rank, m_override=args.states, synth=tuple(args.synthetic))
    if args.strong and world > 1:  # every rank generated the same batch; keep this rank's block of rows
        strong_shard(wl, world, rank, naqs_b200.distributed.shard_bounds)
    table = naqs_b200.DeviceTermTable(wl["xy"], wl["yz"], wl["c"], wl["N"], wl["na"], wl["nb"], device=dev)
    M, K, W = len(wl["states"]), table.K, table.words
    h_states = torch.from_numpy(np.ascontiguousarray(wl["states"]).reshape(M, W).view(np.int64)).pin_memory()
    h_psi = torch.from_numpy(wl["psi"]).pin_memory()
    d_states, d_psi = h_states.to(dev), h_psi.to(dev)
    d_eloc = torch.empty((M, 2), dtype=torch.float64, device=dev)
    h_eloc = torch.empty((M, 2), dtype=torch.float64).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    allreduce_table = world > 1 and naqs_b200.distributed.can_allreduce_table(table, d_psi) and not os.environ.get("NAQS_BENCH_ALLGATHER")
    # multi-GPU exchange: the C-ABI communicator (push kernels over peer memory / NCCL all-gather, csrc/comm.cu) unless
    # NAQS_BENCH_TORCH_DIST=1 asks for the torch.distributed collectives of round 1
    comm = naqs_b200.distributed.Comm(dev) if world > 1 and not os.environ.get("NAQS_BENCH_TORCH_DIST") else None
    if world > 1:
        g_states = torch.empty((world * M, W), dtype=torch.int64, device=dev)
        g_psi = torch.empty(world * M, dtype=torch.complex64, device=dev)
        dense_tbl = naqs_b200.distributed.aligned_dense_table(table) if allreduce_table else None
    dedup = wl.get("dedup_table", False)
    if dedup:  # H2O "with replacement" workload: the lookup table is the distinct keys (fixed across steps)
        uk, first = np.unique(wl["states"], return_index=True)
        t_keys = torch.from_numpy(uk.view(np.int64)).to(dev).reshape(-1, 1)
        t_psi = d_psi[torch.from_numpy(first).to(dev)]

    def step(states, psi, out):
        """One pass of the hot path with device-resident inputs; returns the 5 statistics sums (device)."""
        if comm is not None:
            # (key, psi) of every rank -> lookup table of this rank, one collective entry (naqs_table_exchange)
            comm.exchange(table, states, psi, max_local=M, flags=0x1000 if os.environ.get("NAQS_BENCH_ALLGATHER") else int(os.environ.get("NAQS_BENCH_EXCHANGE_FLAGS", "0"), 0))
        elif world > 1 and allreduce_table:
            # small key space: the direct-address table itself is all-reduced (8 * 2^N bytes, independent of the rank count);
            # psi is a function of the state, so copies of a key on several ranks are identical
            naqs_b200.distributed.allreduce_dense_table(table, states, psi, out=dense_tbl)
        elif world > 1:
            # all-gather (key, psi) -> lookup build (one amplitude per key) -> fused E_loc on the shard -> all-reduce of 5 sums
            g_k, g_p, _ = naqs_b200.distributed.gather_table(states, psi, equal_sizes=True, out=(g_states, g_psi))
            table.build_lookup(g_k, g_p, duplicates_equal=True)
        elif dedup:
            table.build_lookup(t_keys, t_psi, assume_unique=True)
        else:
            table.build_lookup(states, psi, assume_unique=True)  # distinct keys by construction (energy.py:245 contract)
        ev_k0.record()
        table.local_energy(states, psi, out=out, rebuild_lookup=False)
        ev_k1.record()
        if comm is not None:
            return comm.allreduce_stats(table.stats(out))
        return naqs_b200.distributed.reduce_stats(table.stats(out))

    ev_k0, ev_k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----------------------------------------------------------------
    for _ in range(args.warmup):
        flush.fill_(1)
        step(d_states, d_psi, d_eloc)
    barrier()
    step_ms, kern_ms = [], []
    launches0 = naqs_b200.launch_count()
    with ClockSampler(local_rank) as clocks:
        barrier()
        for _ in range(args.steps):
            flush.fill_(1)  # L2 flush between timed iterations (untimed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s5 = step(d_states, d_psi, d_eloc)
            e1.record()
            e1.synchronize()
            step_ms.append(e0.elapsed_time(e1))
            kern_ms.append(ev_k0.elapsed_time(ev_k1))
        barrier()
    launches = naqs_b200.launch_count() - launches0
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = world * M * K * args.steps / (total_ms * 1e-3)
    stats = s5.cpu().numpy()

    # ---- multi-GPU self-check: the all-reduced statistics against ONE rank recomputing the global batch -------------------
    mgpu_check = None
    if world > 1:
        all_k = torch.empty((world * M, W), dtype=torch.int64, device=dev)
        all_p = torch.empty(world * M, dtype=d_psi.dtype, device=dev)
        dist.all_gather_into_tensor(all_k, d_states.contiguous())
        dist.all_gather_into_tensor(torch.view_as_real(all_p), torch.view_as_real(d_psi.contiguous()))
        if rank == 0:
            t1 = naqs_b200.DeviceTermTable(wl["xy"], wl["yz"], wl["c"], wl["N"], wl["na"], wl["nb"], device=dev)
            t1.build_lookup(all_k, all_p, duplicates_equal=True)
            e1 = t1.local_energy(all_k, all_p, rebuild_lookup=False)
            s1 = t1.stats(e1).cpu().numpy()
            mean_m, mean_1 = stats[1] / stats[0], s1[1] / s1[0]
            var_m, var_1 = stats[3] / stats[0] - mean_m ** 2, s1[3] / s1[0] - mean_1 ** 2
            mgpu_check = {"rows": int(s1[4]), "mean_rel_diff": float(abs(mean_m - mean_1) / abs(mean_1)),
                          "var_rel_diff": float(abs(var_m - var_1) / max(abs(var_1), 1e-300)), "n_equal": bool(int(stats[4]) == int(s1[4]))}
            mgpu_check["ok"] = bool(mgpu_check["n_equal"] and mgpu_check["mean_rel_diff"] <= 1e-12 and mgpu_check["var_rel_diff"] <= 1e-9)
            # row-level: this rank's shard of the multi-GPU run against the same rows of the single-rank run
            mgpu_check["shard_max_rel_diff"] = float((d_eloc - e1[:M]).abs().max() / e1[:M].abs().max())
            mgpu_check["ok"] = bool(mgpu_check["ok"] and mgpu_check["shard_max_rel_diff"] <= 1e-12)
            del t1, e1
        del all_k, all_p
        okt = torch.tensor([1 if (mgpu_check is None or mgpu_check["ok"]) else 0], device=dev)
        dist.broadcast(okt, 0)
        if int(okt.item()) != 1:
            if rank == 0:
                print(json.dumps({"error": "multi-GPU statistics differ from the single-rank recomputation", "check": mgpu_check}), file=sys.stderr)
            dist.destroy_process_group()
            return 1

    # ---- end to end through the host-buffer public API -----------------------------------------
    e2e = None
    if not args.no_e2e:
        def e2e_step():
            if world == 1 and not dedup:
                return table.local_energy_host(h_keys_np, h_psi_np, out=h_eloc32_np, assume_unique=True, out_dtype=np.complex64)
            ds, dp = h_states.to(dev, non_blocking=True), h_psi.to(dev, non_blocking=True)
            step(ds, dp, d_eloc)
            h_eloc.copy_(d_eloc, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return h_eloc
        # the reference-facing call: state indices in the reference's index dtype (int32 for 16 <= N < 30, hilbert.py:405-410),
        # psi complex64, E_loc back as float32 pairs (complex.py:139-140); all three buffers page-locked
        idt = torch.int16 if wl["N"] < 16 else (torch.int32 if wl["N"] < 30 else torch.int64)
        if W == 1:
            h_keys = h_states.reshape(-1).to(idt).pin_memory()
            h_keys_np = h_keys.numpy() if idt != torch.int64 else h_keys.numpy().view(np.uint64)
        else:
            h_keys_np = h_states.numpy().view(np.uint64)
        h_psi_np = h_psi.numpy()
        h_eloc32 = torch.empty(M, dtype=torch.complex64).pin_memory()
        h_eloc32_np = h_eloc32.numpy()
        for _ in range(3):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(5, min(args.steps, 50))
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        host_api = world == 1 and not dedup
        serial_value = world * M * K * n_e2e / float(dt.item())
        e2e_value, pipeline = serial_value, None
        if host_api:
            # the same call with THREE batches in flight (naqs_eloc_host_begin / _end, one table handle and one output buffer per
            # batch): the PCIe copies of one batch overlap the kernel of the other.  Every step still uploads its inputs from
            # page-locked host memory and downloads its E_loc inside the timed region.
            depth = 3
            tabs = [table] + [naqs_b200.DeviceTermTable(wl["xy"], wl["yz"], wl["c"], wl["N"], wl["na"], wl["nb"], device=dev) for _ in range(depth - 1)]
            outs = [h_eloc32_np] + [torch.empty(M, dtype=torch.complex64).pin_memory().numpy() for _ in range(depth - 1)]

            def pipelined(n):
                for i in range(n):
                    if i >= depth:
                        tabs[i % depth].local_energy_host_wait()
                    tabs[i % depth].local_energy_host(h_keys_np, h_psi_np, out=outs[i % depth], assume_unique=True, out_dtype=np.complex64, wait=False)
                for t_ in tabs:
                    t_.local_energy_host_wait()
            pipelined(2 * depth)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pipelined(n_e2e)
            torch.cuda.synchronize()
            dt_p = time.perf_counter() - t0
            assert all(np.array_equal(outs[0], o) for o in outs[1:])  # every handle computed the same batch
            e2e_value = M * K * n_e2e / dt_p
            pipeline = {"batches_in_flight": depth, "serial_value": serial_value,
                        "note": "value = steady-state rate with three batches in flight (begin/end form of the same call); serial_value = one "
                                "synchronous call at a time, where upload, kernel and download of a batch cannot overlap (the table is the batch itself)"}
            del tabs[1:]
        e2e = {"value": e2e_value, "unit": UNIT, "pipeline": pipeline,
               "h2d_bytes_per_step": int(M * (h_keys_np.itemsize * (1 if W == 1 else W) + 8)) if host_api else int(M * (8 * W + 8)),
               "d2h_bytes_per_step": int(M * 8) if host_api else int(M * 16), "steps": n_e2e,
               "api": "DeviceTermTable.local_energy_host -> naqs_eloc_host: pinned host state indices (reference index dtype) + complex64 psi in, "
                      "E_loc out as float32 pairs (what calculate_local_energy returns, complex.py:139-140); complex128 arithmetic" if host_api
                      else "pinned H2D of keys + psi, multi-GPU step, pinned D2H of complex128 E_loc"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline + cpu baseline (rank 0) -------------------------------------------------------
    peaks, peak_src = measured_peaks()
    k_ms = float(np.mean(kern_ms))
    T = (world * M) if not dedup else int(t_keys.shape[0])
    # algorithmic HBM bytes of one launch: states + psi in, E_loc out, Pauli table, lookup entries touched once
    algo_bytes = M * (8 * W + 8 + 16) + K * (16 * W + 8) + min(T, 2 ** min(wl["N"], 40)) * 16
    achieved_gbs = algo_bytes / (k_ms * 1e-3) / 1e9
    ncu = ncu_summary().get(args.workload, {})
    traffic = (ncu.get("dram_bytes_read", 0) + ncu.get("dram_bytes_write", 0)) if ncu else None
    roofline_hbm = {"bound": "hbm", "achieved": achieved_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved_gbs / peaks["hbm_gbs"],
                    "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes": int(algo_bytes),
                    "note": "HBM is NOT the binding resource of this path (0.03 B per coupling, SURVEY.md §8d)"}
    # binding resource: L1TEX wavefronts / warp-instruction issue of the formulation's algorithmic work, computed from THIS run's
    # kernel time and SM clock (nothing is read from a committed profile)
    clk_summary = clocks.summary()
    sm_mhz = float(clk_summary.get("sm_mhz") or clk_summary.get("sm_max_mhz") or 1965.0)
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    table_kind = "dense" if wl["N"] <= 22 else "hash"  # NAQS_LOOKUP_AUTO (naqs_lookup_build)
    keyorder_mode = wl["N"] <= 26 and W == 1 and table_kind == "dense" and (T >= (1 << wl["N"]) // 8)
    units = stream_units(np.asarray(wl["xy"]))
    n_units = ((1 << wl["N"]) // 32) if keyorder_mode else (M + 31) // 32
    nn = 5 if wl["N"] <= 20 else (8 if wl["N"] <= 32 else (16 if wl["N"] <= 63 else 32))
    pr = pipe_roofline("keyorder" if keyorder_mode else "hash" if table_kind == "hash" else "dense-rows", units, n_units, nn, k_ms, sm_mhz, sm_count)
    wave_rate = pr["warp_units"] * pr["wavefronts_per_unit"] / (k_ms * 1e-3) / 1e9
    instr_rate = pr["warp_units"] * pr["warp_instr_per_unit"] / (k_ms * 1e-3) / 1e9
    xu_rate = pr["warp_units"] * pr["xu_instr_per_unit"] / (k_ms * 1e-3) / 1e9
    ach, pk, un = {"l1tex": (wave_rate, WAVEFRONTS_PER_CLK_SM, "Gwavefront/s"), "issue": (instr_rate, ISSUE_PER_CLK_SM, "Gwarp-instr/s"),
                   "xu": (xu_rate, XU_PER_CLK_SM, "Gwarp-instr/s (XU pipe)")}[pr["bound"]]
    roofline = {"bound": pr["bound"], "achieved": ach, "peak": pk * sm_count * sm_mhz * 1e-3, "unit": un, "frac": pr["frac"], "traffic": traffic,
                "kernel": ("eloc_keyorder_kernel" if keyorder_mode else "eloc_sliced_kernel") + " (timed with CUDA events around naqs_eloc: includes its mark / bin / finalize helpers)",
                "kernel_ms": k_ms, "model": pr,
                "peak_source": "algorithmic wavefronts and warp instructions of the formulation (bench.py pipe_roofline, DESIGN.md §6) against 1 wavefront, "
                               "4 warp instructions and 0.5 XU (F2F.F64.F32) warp instructions per clock and SM at the SM clock sampled during this run",
                "hbm": roofline_hbm}
    pp, pp_name = pipe_peaks()
    kernel_rate = M * K / (k_ms * 1e-3)
    roofline_pipe = {"kernel_couplings_per_s": kernel_rate, "unit": UNIT}
    if pp:
        roofline_pipe.update({"direct_formulation_ceiling": pp.get("direct_couplings_per_s"),
                              "vs_direct_ceiling": kernel_rate / pp["direct_couplings_per_s"] if pp.get("direct_couplings_per_s") else None,
                              "popc_ops_per_s": pp.get("popc_ops_per_s"), "dadd_ops_per_s": pp.get("dadd_ops_per_s"),
                              "gather16B_L2_per_s": pp.get("gather16B_16MB_per_s"), "peak_source": f"profiles/{pp_name}"})
    if ncu:  # context only: pipe utilisation of the committed ncu capture of this workload (NOT the roofline fraction above)
        roofline_pipe.update({"l1tex_data_pipe_pct_ncu": ncu["l1tex_data_pipe_pct"], "issue_active_pct_ncu": ncu["issue_active_pct"],
                              "ncu_source": ncu.get("source")})
    extras = None
    if world == 1 and not args.no_extras:
        try:
            extras = measure_extras(naqs_b200, dev, args)
        except Exception as e:  # noqa: BLE001
            extras = {"error": repr(e)}
    cpu, parity_vs_cpu = None, None
    if world == 1 and args.cpu_sample > 0:
        try:
            cpu, (c_st, c_psi, c_eloc) = cpu_baseline_sample(wl, args.cpu_sample)
            # the same sample (table = the sample itself, as the reference path computed it) through the device path: the bench
            # line carries its own parity figure against the reference kernels (bar: 1e-12 relative, north_star)
            try:
                g = _lib.complex_from_pairs(table.local_energy(np.asarray(wl["states"])[: len(c_psi)], c_psi, assume_unique=not dedup))
                rel = np.abs(g - np.asarray(c_eloc)) / np.maximum(np.abs(np.asarray(c_eloc)), 1e-300)
                parity_vs_cpu = {"rows": int(len(g)), "max_rel_err": float(rel.max()), "ok": bool(rel.max() <= 1e-12)}
            except Exception as e:  # noqa: BLE001
                parity_vs_cpu = {"error": repr(e)}
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {e}"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": wl["desc"], "states_per_gpu": M, "terms": K, "lookup_keys": T,
                       "parallelism": f"states sharded x{world}, Pauli table replicated" + (
                           ((", naqs_table_exchange: the ranks' 2^N-entry complex64 tables merged over peer memory (two-shot: P2P loads of one slice, P2P stores of the merged slice; push kernel for sparse shards)" if allreduce_table and not os.environ.get("NAQS_BENCH_ALLGATHER") else ", naqs_table_exchange: NCCL all-gather of (key, psi) + lookup build")
                            + " + naqs_stats_allreduce of 5 fp64 sums") if comm is not None else
                           ((", NCCL all-reduce (MAX) of the 2^N-entry complex64 amplitude table" if allreduce_table else ", NCCL all-gather of (key, psi)") + " + all-reduce of 5 fp64 sums" if world > 1 else "")),
                       "l2": "256 MB device memset between timed steps (L2 flush, untimed)", "timing": "CUDA events per step on the launching stream, max over ranks"},
            "clocks": clk_summary, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "roofline_pipe": roofline_pipe,
            "cpu_baseline": cpu, "other_configs": extras,
            "check": {"mean_eloc_re": float(stats[1] / stats[0]), "n": int(stats[4]), "multi_gpu_vs_single_rank": mgpu_check,
                      "vs_cpu_baseline_sample": parity_vs_cpu}}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
