"""Regenerate profiles/README.md from the JSON / CSV evidence in profiles/ (round tag as argv[1], default r01)."""
import json
import os
import sys

R = sys.argv[1] if len(sys.argv) > 1 else "r01"
P = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")


def last_json(name):
    with open(os.path.join(P, name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


b = last_json(f"{R}_final_bench_n2_1e6.json")
ref = last_json(f"{R}_final_bench_reference.json")
ncu = json.load(open(os.path.join(P, f"ncu_summary_{R}.json")))
oc = b.get("other_configs", {})
out = []
w = out.append
w(f"# profiles/ — round {int(R[1:])} evidence (1 x B200, {b['clocks']['sm_mhz']:.0f} MHz, throttle reasons {b['clocks']['reasons']}; `{R}_final_smi.csv`)\n")
w("Regenerate this file with `python bench_tools/make_profiles_readme.py`.\n")
w("## Bench lines (`python bench.py`, `python bench.py --impl reference`)\n")
w("| arm | workload | value (couplings/s) | ms/step | e2e (couplings/s) |\n|---|---|---|---|---|")
w(f"| B200 (`{R}_final_bench_n2_1e6.json`) | N2 STO-3G, M = 1e6, K = 2239 | {b['value']:.3e} | {b['ms_per_step']:.4f} | {b['e2e']['value']:.3e} |")
w(f"| reference CPU path, {ref['cpu_baseline']['cores']} host cores (`{R}_final_bench_reference.json`) | same, {ref['cpu_baseline']['sample'].split(';')[0]} | {ref['value']:.3e} | {ref['ms_per_step']:.1f} | = value |\n")
if oc:
    h, l, v = oc.get("h2o_1e5", {}), oc.get("li2o_1e5", {}), oc.get("lih_vmc_eloc_call", {})
    w(f"Other BASELINE configs from the same run (`other_configs`): H2O 1e5 rows {h.get('value', 0):.3e} ({1e3 * h.get('ms_per_step', 0):.0f} us/step), "
      f"Li2O 1e5 sector states {l.get('value', 0):.3e} ({l.get('ms_per_step', 0):.3f} ms/step); LiH VMC E_loc call ({v.get('states')} states): "
      f"B200 host call {1e3 * v.get('b200_host_call_ms', 0):.0f} us vs reference CPU {1e3 * v.get('reference_cpu_cold_cache_ms', 0):.0f} us (cold H cache) / "
      f"{1e3 * v.get('reference_cpu_warm_cache_ms', 0):.0f} us (warm)."
      + (f" CSR rows of the full N2 sector ({oc['n2_sector_csr_rows'].get('nnz')} stored elements): {oc['n2_sector_csr_rows'].get('b200_rows_ms', 0):.2f} ms vs reference update_H "
         f"{oc['n2_sector_csr_rows'].get('reference_cpu_update_H_ms', 0):.0f} ms." if "n2_sector_csr_rows" in oc and "nnz" in oc["n2_sector_csr_rows"] else "") + "\n")
n, l = ncu["n2_1e6"], ncu["li2o_1e5"]
w(f"## Hot kernel, `ncu --set full --clock-control none` (`{R}_final_n2_ncu_raw.csv`, `{R}_final_li2o_ncu_raw.csv`; summary `ncu_summary_{R}.json`)\n")
w("| metric | N2 1e6 (key-order walk, dense complex64 table) | Li2O 1e5 (hash lookup, Bloom filter in shared memory, survivor queue) |\n|---|---|---|")
rows = [("kernel", lambda d: f"`{d['kernel']}`"), ("grid x block, dynamic smem", lambda d: f"{d['grid']} x {d['block']}, {float(d['smem_dynamic_kb']):.1f} KB"),
        ("duration", lambda d: f"{d['duration_us']:.1f} us"), ("registers / thread", lambda d: f"{d['registers']:.0f}"),
        ("L1TEX data-pipe wavefronts (% of peak)", lambda d: f"**{d['l1tex_data_pipe_pct']:.1f}**"),
        ("warp-instruction issue (% of peak)", lambda d: f"{d['issue_active_pct']:.1f}"),
        ("ALU / FP64 / XU pipe (% of peak)", lambda d: f"{d['alu_pipe_pct']:.1f} / {d['fp64_pipe_pct']:.1f} / {d['xu_pipe_pct']:.1f}"),
        ("shared-memory wavefronts (of which bank conflicts)", lambda d: f"{d['shared_wavefronts']:.3e} ({d['shared_bank_conflict_wavefronts']:.2e})"),
        ("L2 throughput (% of peak)", lambda d: f"{d['lts_throughput_pct']:.1f}"),
        ("DRAM bytes read + written", lambda d: f"{d['dram_bytes_read'] / 1e6:.1f} MB + {d['dram_bytes_write'] / 1e3:.1f} KB"),
        ("warp instructions", lambda d: f"{d['warp_instructions']:.3e}")]
for name, fn in rows:
    w(f"| {name} | {fn(n)} | {fn(l)} |")
w(f"\nReading: HBM is idle ({n['dram_bytes_read'] / 1e6:.1f} MB per launch against {b['roofline']['algorithmic_bytes'] / 1e6:.0f} MB of algorithmic bytes — the table and the batch sit in L2), as "
  "SURVEY.md §8d predicted; both kernels are bound by the L1TEX wavefront pipe (shared-memory LUT / parity-table / filter reads + table gathers). "
  f"The direct AND/POPC formulation tops out at 4.02e12 couplings/s (POPC pipe, `pipe_peaks_{R}.json`); the sliced kernel runs N2 at "
  f"{b['roofline_pipe']['kernel_couplings_per_s']:.2e} kernel-only.\n")
w(f"## Launch list (`{R}_final_launches_n2_1e6.md`): see the share of the hot kernel there.\n")
w(f"## Scaling (`{R}_scale_n2_g*.json`; 8-GPU box, `bench.py --gpus N` under torchrun, weak: 10^6 rows per GPU)\n")
w("Each step: every rank scatters its (key, psi) pairs into a 2^20-entry complex64 table, the tables are all-reduced (MAX on bit patterns, 8 MB whatever N), "
  "fused kernel on the rank's rows, all-reduce of 5 fp64 sums.\n")
w("| GPUs | value (couplings/s) | ms/step | kernel ms | vs N x 1-GPU |\n|---|---|---|---|---|")
base = None
for g in (1, 2, 4, 8):
    fn = f"{R}_scale_n2_g{g}.json"
    if os.path.exists(os.path.join(P, fn)):
        d = last_json(fn)
        base = base or d["value"]
        w(f"| {g} | {d['value']:.3e} | {d['ms_per_step']:.3f} | {d['roofline']['kernel_ms']:.3f} | {100 * d['value'] / (g * base):.0f}% |")
for tag, title in (("scale_li2o", "Li2O weak scaling (10^5 sector states per GPU, hash lookup, all-gather exchange; the lookup table grows with the rank count)"),
                   ("strong_li2o", "Li2O strong scaling (`bench.py --strong`: ONE batch of 10^5 states split over the ranks)")):
    rows_ = []
    for g in (1, 2, 4, 8):
        fn = f"{R}_{tag}_g{g}.json"
        if os.path.exists(os.path.join(P, fn)):
            d = last_json(fn)
            rows_.append(f"| {g} | {d['value']:.3e} | {d['ms_per_step']:.3f} | {d['roofline']['kernel_ms']:.3f} |")
    if rows_:
        w(f"\n{title} (`{R}_{tag}_g*.json`):\n")
        w("| GPUs | value (couplings/s) | ms/step | kernel ms |\n|---|---|---|---|")
        out.extend(rows_)
w(f"\n`{R}_collective_probe_g2.json`: NCCL all-reduce latencies at this size against torch's symmetric-memory kernels (why the exchange stays on NCCL).\n")
rl = os.path.join(P, f"{R}_bench_n2_1e6_rows_leg.json")
if os.path.exists(rl):
    r_ = last_json(f"{R}_bench_n2_1e6_rows_leg.json")["other_configs"].get("n2_sector_csr_rows", {})
    if "nnz" in r_:
        w(f"## CSR-rows mode (`{R}_bench_n2_1e6_rows_leg.json`, `other_configs.n2_sector_csr_rows`)\n")
        w(f"Full N2 sector, {r_['states']} states -> {r_['nnz']} stored matrix elements (same count as the reference): device rows "
          f"(count + scan + fill with restricted column indices) {r_['b200_rows_ms']:.2f} ms vs reference `update_H` {r_['reference_cpu_update_H_ms']:.0f} ms on {r_['cores']} cores.\n")
w("## Other files")
w(f"`pipe_peaks_{R}.json` — measured POPC / LOP3 / IMAD / DADD / gather ceilings (`bench_tools/pipe_peaks.cu`);\n`{R}_sanitize_*.log` — compute-sanitizer memcheck + racecheck: 0 errors, 0 hazards;\n"
  f"`{R}_v1_*` — the first (direct, POPC-bound) kernel for comparison: 1.59e12 couplings/s, 1.37 ms kernel.\n")
sw = os.path.join(P, f"{R}_synthetic_sweep.jsonl")
if os.path.exists(sw):
    w(f"## Synthetic Pauli-sum sweep (BASELINE config 5; `{R}_synthetic_sweep.jsonl`, `bench.py --workload synthetic --synthetic N K M`)\n")
    w("Random 2-/4-flip masks (Kxy ~ K/6), JW-like + random Z strings, normal fp64 coefficients, weight-N/2 keys; the lookup table is the batch itself. "
      "Unlike a molecule, random coefficients never cancel, so every one of the K/6 groups of a state is a live coupling whose coupled state misses the table — "
      "this sweep stresses the lookup path (Bloom filter in shared memory for M <= 2^17, in L2 above).\n")
    w("| N (mask width) | K | M | couplings/s | ms/step |\n|---|---|---|---|---|")
    import re
    for line in open(sw):
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        m = re.search(r"N=(\d+) qubits, K=(\d+) terms.*M=(\d+)", d["config"]["workload"])
        N, K, M = (int(x) for x in m.groups())
        w(f"| {N} ({'128' if N > 63 else '64'}-bit) | {K} | {M} | {d['value']:.3e} | {d['ms_per_step']:.3f} |")
open(os.path.join(P, "README.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out)[:3000])
