"""Summarise an `ncu --page source --csv` dump: instruction mix and stall samples per opcode."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ix = {h: i for i, h in enumerate(hdr)}
ops, samples = collections.Counter(), collections.Counter()
tot = tots = 0
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    src = r[ix["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = (m.group(2) if m else src).split(".")[0]
    try:
        n, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    except ValueError:
        continue
    ops[op] += n
    samples[op] += s
    tot += n
    tots += s
print("total warp-instructions", tot, "samples", tots)
for k, v in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{k:10s} {v:12d} {100 * v / tot:5.1f}%   stall-samples {100 * samples[k] / max(1, tots):5.1f}%")
