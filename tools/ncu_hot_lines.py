"""Hot source lines of an .ncu-rep captured with --import-source on (-lineinfo build): warp-state samples per CUDA source line.
usage: python tools/ncu_hot_lines.py report.ncu-rep [N]"""
import csv, io, subprocess, sys
from collections import Counter

rep = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
h = rows[hi]
ix = {}
for i, k in enumerate(h):
    ix.setdefault(k, i)
def I(v):
    try: return int(v)
    except Exception: return 0
data = [r for r in rows[hi + 1:] if len(r) >= len(h)]
tot = sum(I(r[ix["# Samples"]]) for r in data)
c, ins, txt = Counter(), Counter(), {}
for r in data:
    c[r[0]] += I(r[ix["# Samples"]]); ins[r[0]] += I(r[ix["Instructions Executed"]]); txt[r[0]] = r[1]
print("total samples", tot)
for k, v in c.most_common(N):
    print("%6s %6d %5.1f%% inst %9d | %s" % (k, v, 100.0 * v / max(tot, 1), ins[k], txt[k].strip()[:120]))
