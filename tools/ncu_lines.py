"""Per-source-line stall samples / instruction counts from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
python tools/ncu_lines.py file.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, out = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0] != "":
        try:
            line, samples, inst = int(r[0]), int(r[6]), int(r[7])
            lsb = int(r[hdr.index("stall_long_sb")])
        except ValueError:
            continue
        out.append((cur, line, r[1][:100], samples, inst, lsb))
ts, ti = sum(o[3] for o in out), sum(o[4] for o in out)
print("total samples", ts, "warp instructions", ti)
for o in sorted(out, key=lambda o: -o[3])[:top]:
    print(f"{o[0]}:{o[1]:4d} samp {100 * o[3] / ts:5.1f}% inst {100 * o[4] / ti:5.1f}% long_sb {100 * o[5] / ts:5.1f}% | {o[2]}")
