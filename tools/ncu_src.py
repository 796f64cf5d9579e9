"""Summarise an ncu report's per-instruction stall samples: python tools/ncu_src.py rep.ncu-rep <launch-skip> [top]"""
import csv, subprocess, sys, io
from collections import Counter
rep, skip = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if "Source" in r)
hdr = rows[h]
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
body = [r for r in rows[h + 1:] if len(r) > max(ia, isamp, iex)]
# the page lists the function twice in some versions: keep the first copy
seen = set(); data = []
for i, r in enumerate(body):
    if r[0] in seen: break
    seen.add(r[0])
    try: data.append((int(r[isamp] or 0), int(r[iex] or 0), r[ia].strip(), i))
    except ValueError: pass
tot = sum(d[0] for d in data)
print("kernel:", rows[0][1] if len(rows[0]) > 1 else "?", "| samples", tot, "| sass lines", len(data))
c = Counter(); e = Counter()
for s, x, src, i in data:
    t = src.split(); op = t[1] if t[0].startswith("@") else t[0]
    c[op] += s; e[op] += x
for op, s in c.most_common(top): print(f"{op:30s} {s:7d} {100*s/max(tot,1):5.1f}%  executed {e[op]}")
print("--- top lines")
for s, x, src, i in sorted(data, reverse=True)[:top]: print(i, s, x, src)
