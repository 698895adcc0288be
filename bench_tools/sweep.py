"""BASELINE config 5 — synthetic Pauli-sum sweep (1e3..1e6 terms x 1e4..1e7 states, 64- and 128-bit masks) in ONE process.

    python bench_tools/sweep.py [--out gpurun_out/sweep.jsonl] [--widths 63 127] [--terms 1000 10000 100000 1000000]
                                [--states 10000 100000 1000000 10000000]

Per grid point one JSON line: device-resident couplings/s (lookup build + fused kernel per step, CUDA events, 256 MB L2 flush
between steps), the kernel time alone, and the live pipe roofline of that point (bench.pipe_roofline: algorithmic L1TEX
wavefronts / warp instructions of the hash-walk formulation against 1 wavefront and 4 instructions per clock and SM).
The lookup table is the batch itself; random coefficients never cancel, so every group of every state is a live coupling
whose coupled state misses the table — the sweep stresses the filter / lookup path rather than the exact-zero skip."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import naqs_b200  # noqa: E402


def random_states(N, M, seed):
    """M distinct random keys of N bits ([M, W] uint64); exact weight N/2 for small batches (bench.make_workload), plain random
    bits above 2e6 states (the per-row argsort of the exact-weight generator needs 8 * N * M bytes)."""
    W = 1 if N <= 63 else 2
    if M <= 2_000_000:
        return np.ascontiguousarray(bench.make_workload("synthetic", seed, synth=(N, 1000, M))["states"]).reshape(-1, W)
    rng = np.random.default_rng(seed)
    st = np.zeros((int(M * 1.01) + 16, W), np.uint64)
    for w in range(W):
        bits = min(64, N - 64 * w)
        st[:, w] = rng.integers(0, 2 ** min(bits, 63), size=len(st), dtype=np.int64).astype(np.uint64)
        if bits == 64:
            st[:, w] |= rng.integers(0, 2, size=len(st), dtype=np.int64).astype(np.uint64) << np.uint64(63)
    st = np.unique(st, axis=0)
    return st[rng.permutation(len(st))[:M]]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    ap.add_argument("--widths", type=int, nargs="+", default=[63, 127])
    ap.add_argument("--terms", type=int, nargs="+", default=[1000, 10000, 100000, 1000000])
    ap.add_argument("--states", type=int, nargs="+", default=[10000, 100000, 1000000, 10000000])
    ap.add_argument("--budget-s", type=float, default=2.0, help="timed seconds per grid point (at least 3 steps)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        for N in args.widths:
            W = 1 if N <= 63 else 2
            nn = 16 if N <= 63 else 32
            states_of = {}
            for K in args.terms:
                t0 = time.time()
                xy, yz, c = bench.synthetic_table(N, K)
                table = naqs_b200.DeviceTermTable(xy, yz, c, N, None, None, device=dev)
                units = bench.stream_units(xy)
                t_table = time.time() - t0
                for M in args.states:
                    if M not in states_of:
                        states_of[M] = random_states(N, M, seed=1)
                    st = states_of[M]
                    M_eff = len(st)
                    d_states = torch.from_numpy(st.view(np.int64)).to(dev)
                    d_psi = torch.from_numpy(bench.psi_for(M_eff, 2)).to(dev)
                    out = torch.empty((M_eff, 2), dtype=torch.float64, device=dev)
                    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

                    def step():
                        ev[0].record()
                        table.build_lookup(d_states, d_psi, assume_unique=True)
                        ev[1].record()
                        table.local_energy(d_states, d_psi, out=out, rebuild_lookup=False)
                        ev[2].record()

                    for _ in range(3):
                        flush.fill_(1)
                        step()
                    torch.cuda.synchronize()
                    est = ev[0].elapsed_time(ev[2]) * 1e-3
                    steps = int(max(3, min(50, args.budget_s / max(est, 1e-6))))
                    step_ms, kern_ms = [], []
                    with bench.ClockSampler(0, period=0.005) as clocks:
                        for _ in range(steps):
                            flush.fill_(1)
                            step()
                            ev[2].synchronize()
                            step_ms.append(ev[0].elapsed_time(ev[2]))
                            kern_ms.append(ev[1].elapsed_time(ev[2]))
                    k_ms = float(np.mean(kern_ms))
                    sm_mhz = float(clocks.summary().get("sm_mhz") or 1965.0)
                    pr = bench.pipe_roofline("hash", units, (M_eff + 31) // 32, nn, k_ms, sm_mhz, sm_count)
                    line = {"workload": f"synthetic N={N} ({64 * W}-bit masks) K={K} M={M_eff}", "N": N, "mask_bits": 64 * W, "K": K, "Kxy": int(len(np.unique(xy, axis=0))),
                            "M": M_eff, "steps": steps, "value": M_eff * K / (np.mean(step_ms) * 1e-3), "unit": bench.UNIT, "ms_per_step": float(np.mean(step_ms)),
                            "kernel_ms": k_ms, "kernel_couplings_per_s": M_eff * K / (k_ms * 1e-3),
                            "roofline": {"bound": pr["bound"], "frac": pr["frac"], "t_roof_ms": pr["t_roof_ms"], "t_issue_ms": pr["t_issue_ms"], "t_l1tex_ms": pr["t_l1tex_ms"]},
                            "sm_mhz": sm_mhz, "table_build_s": t_table,
                            "mean_eloc_re": float(out[:, 0].mean().item())}
                    f.write(json.dumps(line) + "\n")
                    f.flush()
                    print("N %3d K %7d M %8d : %.3e couplings/s  %.3f ms/step  kernel %.3f ms  roofline frac %.3f (%s)" % (
                        N, K, M_eff, line["value"], line["ms_per_step"], k_ms, pr["frac"], pr["bound"]), flush=True)
                    del d_states, d_psi, out
                del table
                torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
