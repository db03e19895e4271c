"""Rank CUDA source lines of an ncu `--page source --print-source cuda,sass --csv` export by stall
samples / executed warp instructions.  usage: python profiles/rank_lines.py export.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None; data = {}
for r in rows:
    if r and r[0] == "Line No": hdr = r; continue
    if r and r[0] == "File Path" and hdr is not None and data: break      # first kernel instance only
    if hdr is None or len(r) < 10 or r[2] != "-": continue                 # keep source-line summary rows
    try:
        ln = int(r[0]); s = int(r[hdr.index("# Samples")]); i = int(r[hdr.index("Instructions Executed")])
    except ValueError:
        continue
    data[ln] = (r[1], s, i)
ts = sum(v[1] for v in data.values()); ti = sum(v[2] for v in data.values())
print("total samples %d, warp instructions %d" % (ts, ti))
for ln, (src, s, i) in sorted(data.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%4d samp %5.1f%% inst %5.1f%% (%8d) %s" % (ln, 100 * s / ts, 100 * i / ti, i, src.strip()[:100]))
