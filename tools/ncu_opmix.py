"""Summarise an `ncu --page source --csv` dump: executed warp-instructions per SASS opcode and the
top stall lines.  usage: ncu -i X.ncu-rep --page source --csv | python tools/ncu_opmix.py"""
import collections, csv, re, sys
rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
isrc, iex, ist = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
tot, stall = collections.Counter(), []
for r in rows[hi + 1:]:
    try:
        n = int(r[iex])
    except (ValueError, IndexError):
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
    op = m.group(2) if m else r[isrc][:20]
    parts = op.split(".")
    key = parts[0] + ("." + parts[1] if parts[0] in ("IDP", "LDG", "LDS", "STS", "F2F", "F2I", "I2F", "I2FP", "MUFU", "ATOMS", "SHF") and len(parts) > 1 else "")
    tot[key] += n
    try:
        stall.append((int(r[ist]), r[isrc].strip()[:90], n))
    except ValueError:
        pass
s = sum(tot.values())
print(f"total warp-instructions executed: {s}")
for k, v in tot.most_common(40):
    print(f"  {k:14s} {v:12d} {100 * v / s:5.1f}%")
print("top stall-sample lines:")
ts = sum(x[0] for x in stall)
for st, src, n in sorted(stall, reverse=True)[:25]:
    print(f"  {100 * st / ts:5.1f}%  exec={n:9d}  {src}")
