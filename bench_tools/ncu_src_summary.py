"""Summarise an `ncu --page source --csv` dump: opcode mix, stall samples, hottest SASS lines."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr, data = rows[hi], rows[hi + 1:]
ix = {h: i for i, h in enumerate(hdr)}
tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
tot_samp = sum(int(r[ix["# Samples"]]) for r in data)
print("total warp instr", tot_inst, "samples", tot_samp, "n sass lines", len(data))
ops, samp = collections.Counter(), collections.Counter()
for r in data:
    src = r[ix["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2).split(".")[0] if m else src[:10]
    ops[op] += int(r[ix["Instructions Executed"]])
    samp[op] += int(r[ix["# Samples"]])
print("top opcodes by executed count:")
for op, c in ops.most_common(24):
    print(f"  {op:12s} {c:12d} {100*c/tot_inst:5.1f}%   samples {100*samp[op]/tot_samp:5.1f}%")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("hottest SASS lines by samples:")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:n]:
    print(f"  {int(r[ix['# Samples']]):6d} {100*int(r[ix['# Samples']])/tot_samp:4.1f}%  exec {int(r[ix['Instructions Executed']]):10d} thr {r[ix['Avg. Threads Executed']]:>5s}  {r[ix['Source']].strip()[:100]}")
