"""profiles/ncu_summary_rNN.json from `ncu --page raw --csv` exports of the hot kernel.

    ncu -i gpurun_out/final_n2.ncu-rep --page raw --csv > profiles/r01_final_n2_ncu_raw.csv
    python bench_tools/ncu_summarize.py profiles/ncu_summary_r01.json n2_1e6=profiles/r01_final_n2_ncu_raw.csv li2o_1e5=profiles/...

bench.py reads the summary for roofline.traffic (DRAM bytes per launch) and roofline_pipe (which on-chip resource binds)."""
import csv
import json
import sys


def num(x):
    try:
        return float(str(x).replace(",", ""))
    except ValueError:
        return None


def summarize(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    row = next(r for r in rows[2:] if any(("eloc_sliced_kernel" in c or "eloc_keyorder_kernel" in c) for c in r[:12]))
    d, u = dict(zip(hdr, row)), dict(zip(hdr, units))

    def bytes_of(key):
        v, unit = num(d.get(key)), (u.get(key) or "").lower()
        if v is None:
            return None
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)

    dur, dunit = num(d["gpu__time_duration.sum"]), (u.get("gpu__time_duration.sum") or "").lower()
    dur_us = dur * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(dunit, 1.0)
    return {
        "kernel": d["Kernel Name"].split("(")[0],
        "duration_us": dur_us,
        "dram_bytes_read": bytes_of("dram__bytes_read.sum"),
        "dram_bytes_write": bytes_of("dram__bytes_write.sum"),
        "l1tex_data_pipe_pct": num(d.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed")),
        "issue_active_pct": 100.0 * num(d.get("smsp__issue_active.avg.per_cycle_active")),
        "warp_instructions": num(d.get("smsp__inst_executed.sum")),
        "registers": num(d.get("launch__registers_per_thread")),
        "lts_throughput_pct": num(d.get("lts__throughput.avg.pct_of_peak_sustained_elapsed")),
        "alu_pipe_pct": num(d.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active")),
        "fp64_pipe_pct": num(d.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")),
        "xu_pipe_pct": num(d.get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active")),
        "shared_wavefronts": num(d.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")),
        "shared_bank_conflict_wavefronts": num(d.get("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")),
        "grid": d.get("launch__grid_size"),
        "block": d.get("launch__block_size"),
        "smem_dynamic_kb": d.get("launch__shared_mem_per_block_dynamic"),
        "source": path,
    }


if __name__ == "__main__":
    out = {}
    for arg in sys.argv[2:]:
        name, path = arg.split("=", 1)
        out[name] = summarize(path)
    with open(sys.argv[1], "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))
