"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line."""
import csv
import sys

path, topn = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur_file, hdr, items, tot_i, tot_s = None, None, [], 0, 0
for r in csv.reader(open(path)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= ie or not r[0].isdigit():
        continue
    try:
        n, s = int(r[ie]), int(r[ss])
    except ValueError:
        continue
    tot_i += n
    tot_s += s
    items.append((s, n, cur_file, r[0], r[1][:120]))
items.sort(reverse=True)
print(f"total warp-inst {tot_i}  samples {tot_s}")
for s, n, f, l, t in items[:topn]:
    print(f"{100*s/max(tot_s,1):5.1f}% stall-samples {100*n/max(tot_i,1):5.1f}% inst  {f}:{l}  {t}")
