#!/usr/bin/env python
"""Summary of an `ncu --page source --csv` export: warp-stall samples by reason, executed warp-instructions by opcode, and the
instructions that collect the most stall samples.  usage: ncu_source_summary.py source.csv [cells_per_launch]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
head = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
names, data = rows[head], rows[head + 1:]
col = {n: i for i, n in enumerate(names)}
reasons = [n for n in names if n.startswith("stall_") and "Not Issued" not in n]
by_reason, by_opcode, samples_by_opcode = collections.Counter(), collections.Counter(), collections.Counter()
top = []
for r in data:
    try:
        samples, executed = int(r[col["# Samples"]]), int(r[col["Instructions Executed"]])
    except (ValueError, IndexError):
        continue
    text = re.sub(r"^@!?U?P\d+\s+", "", r[col["Source"]].strip())
    opcode = text.split()[0].split(".")[0] if text else "?"
    by_opcode[opcode] += executed
    samples_by_opcode[opcode] += samples
    for n in reasons:
        if r[col[n]] not in ("", "0"):
            by_reason[n] += int(r[col[n]])
    top.append((samples, executed, r[col["Source"]].strip()))
total_samples, total_instr = sum(by_reason.values()), sum(by_opcode.values())
print(f"warp-stall samples: {total_samples}; executed warp-instructions: {total_instr}")
if len(sys.argv) > 2:
    print(f"thread-instructions per cell: {total_instr * 32 / float(sys.argv[2]):.3f}")
print("\nstall reason            samples   share")
for n, v in by_reason.most_common():
    print(f"{n:24s}{v:8d}  {100.0 * v / total_samples:5.1f} %")
print("\nopcode      warp-instructions   share   stall samples")
for n, v in by_opcode.most_common(16):
    print(f"{n:12s}{v:16d}  {100.0 * v / total_instr:5.1f} %  {samples_by_opcode[n]:8d}")
print("\ninstructions with the most stall samples")
for samples, executed, text in sorted(top, reverse=True)[:12]:
    print(f"{samples:7d}  {text}")
