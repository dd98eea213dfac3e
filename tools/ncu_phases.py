"""Per-phase instruction / stall-sample summary of an `ncu --page source --csv` dump of a fused STFT
kernel: phases are delimited by the CTA barriers (BAR.SYNC) that execute once per tile."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
frames = float(sys.argv[2])
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ix = {k: i for i, k in enumerate(hdr)}
ins = []
for r in rows[h + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        ex, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    except ValueError:
        continue
    ins.append((r[ix["Source"]].strip(), ex, s))
tot = sum(e for _, e, _ in ins)
tots = sum(s for _, _, s in ins)
print(f"total warp-instr/frame {tot / frames:.1f}  samples {tots}")
bounds = [i for i, (src, ex, _) in enumerate(ins) if "BAR.SYNC" in src and ex / frames > 0.01]
bounds = [0] + bounds + [len(ins)]
for a, b in zip(bounds[:-1], bounds[1:]):
    ops = collections.Counter()
    n = s = 0
    for src, ex, sm in ins[a:b]:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
        ops[(m.group(2) if m else src).split(".")[0]] += ex / frames
        n += ex
        s += sm
    print(f"[{a}:{b}] instr/frame {n / frames:.1f}  samples {100 * s / tots:.1f}%")
    print("    " + ", ".join(f"{k}:{v:.1f}" for k, v in ops.most_common(14)))
if len(sys.argv) > 3:
    with open(sys.argv[3], "w") as f:
        for i, (src, ex, s) in enumerate(ins):
            f.write(f"{i:5d} {ex / frames:7.3f} {s:6d}  {src}\n")
