"""Summarise an ncu `--page source --csv` dump: sample share per region between barriers + top SASS lines."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; body = [r for r in rows[2:] if len(r) >= len(rows[1]) - 2 and r[0].startswith("0x")]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
# regions split at BAR.SYNC
reg, cur, start = [], 0, 0
regs = []
acc = {h: 0 for h in stall_cols}
for k, r in enumerate(body):
    s = int(r[isamp] or 0); cur += s
    for i in stall_cols:
        acc[i] += int(r[i] or 0)
    if "BAR.SYNC" in r[isrc] or k == len(body) - 1:
        top = sorted(((v, hdr[i]) for i, v in acc.items() if v), reverse=True)[:3]
        regs.append((start, k, cur, top)); start = k + 1; cur = 0; acc = {h: 0 for h in stall_cols}
for a, b, s, top in regs:
    if s > tot * 0.01:
        print(f"  region sass[{a:5d}..{b:5d}] samples {s:7d} {100*s/tot:5.1f}%  ex={body[a][iex]:>8s}  {top}")
print("top instructions:")
for r in sorted(body, key=lambda r: -int(r[isamp] or 0))[:25]:
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"  {int(r[isamp]):6d} {100*int(r[isamp])/tot:5.1f}%  #{body.index(r):5d} ex={r[iex]:>9s} {r[isrc].strip()[:70]:70s} {st}")
