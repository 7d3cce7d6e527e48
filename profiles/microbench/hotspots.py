"""Summarise an `ncu --page source --csv` export: stall mix and the SASS regions that collect the samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
names = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        data.append((int(r[isamp]), r[isrc][:90], int(r[iex]), {n: int(r[hdr.index(n)] or 0) for n in names}))
    except ValueError:
        pass
data = data[:len(data) // 2] if len(data) > 1 and data[0][1] == data[len(data) // 2][1] else data
tot = sum(d[0] for d in data)
print(rows[0][1][:80], "| samples", tot, "| SASS", len(data), "| warp-instr executed", sum(d[2] for d in data))
agg = {}
for d in data:
    for n, v in d[3].items():
        agg[n] = agg.get(n, 0) + v
print("stall mix:", [(k, v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]])
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 16
for i in range(0, len(data), blk):
    b = data[i:i + blk]
    s, ex = sum(x[0] for x in b), sum(x[2] for x in b)
    if s >= 0.025 * tot:
        top = max(b, key=lambda x: x[0])
        print(f"[{i:4d}-{i + blk - 1:4d}] {100 * s / tot:5.1f}% samples, exec {ex:8d} | top: {top[1].strip()} ({top[0]})")
