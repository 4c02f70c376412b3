"""Merge `ncu -i X.ncu-rep --page raw --csv --print-units base` exports of several reports into one CSV, columns matched BY NAME
(reports of different kernels carry different metric sets).  usage: merge_ncu_csv.py out.csv in1.csv in2.csv ..."""
import csv
import sys

out, files = sys.argv[1], sys.argv[2:]
tabs = []
for f in files:
    r = list(csv.reader(open(f)))
    tabs.append((r[0], r[1], r[2:]))
cols, units = [], {}
for hdr, un, _ in tabs:
    for h, u in zip(hdr, un):
        if h not in units:
            cols.append(h); units[h] = u
        elif u and units[h] and units[h] != u:
            raise SystemExit(f"unit mismatch for {h}: {units[h]} vs {u} (export with --print-units base)")
rows = [cols, [units[c] for c in cols]]
for hdr, _, data in tabs:
    idx = {h: i for i, h in enumerate(hdr)}
    rows += [[d[idx[c]] if c in idx else "" for c in cols] for d in data]
csv.writer(open(out, "w")).writerows(rows)
print(out, len(rows) - 2, "kernels,", len(cols), "columns")
