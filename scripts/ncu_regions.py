"""Per-region (split at BAR.SYNC) instruction and sample shares for one kernel of an ncu report."""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; body = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x")]
# keep the first kernel instance only
first = body[0][0]
idx = [i for i, r in enumerate(body) if r[0] == first]
if len(idx) > 1: body = body[:idx[1]]
iex, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iex]) for r in body); ts = sum(int(r[isamp]) for r in body)
print(f"{kern}: warp-instr {tot}, samples {ts}, sass {len(body)}")
start = 0; cur = 0; cs = 0; acc = {}
for k, r in enumerate(body):
    cur += int(r[iex]); cs += int(r[isamp])
    for i in stall: acc[i] = acc.get(i, 0) + int(r[i] or 0)
    if "BAR.SYNC" in r[isrc] or k == len(body) - 1:
        if cs > ts * 0.01 or cur > tot * 0.01:
            top = sorted(((v, hdr[i][6:]) for i, v in acc.items() if v), reverse=True)[:3]
            print(f"  sass[{start:5d}..{k:5d}] instr {100*cur/tot:5.1f}%  samples {100*cs/ts:5.1f}%  {top}")
        start = k + 1; cur = 0; cs = 0; acc = {}
if len(sys.argv) > 3:
    print("top:")
    for r in sorted(body, key=lambda r: -int(r[isamp]))[:int(sys.argv[3])]:
        st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall), reverse=True)[:2]
        print(f"  {100*int(r[isamp])/ts:5.1f}% #{body.index(r):5d} ex={r[iex]:>9s} {r[isrc].strip()[:64]:64s} {st}")
