"""Summarises an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
data = []
for r in rows:
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        ie, ss, te, st = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr is None or len(r) <= ie or r[0] == "":
        continue
    try:
        data.append((int(r[ie]), int(r[ss]), int(r[te]), r[0], r[1][:120]))
    except ValueError:
        pass
tot = sum(d[0] for d in data)
tots = sum(d[1] for d in data)
print("total warp instructions %d, samples %d" % (tot, tots))
for d in sorted(data, reverse=True)[:top]:
    print("%5.1f%% inst %5.1f%% samples  lanes %4.1f  L%-4s %s" % (100.0 * d[0] / tot, 100.0 * d[1] / max(tots, 1), d[2] / max(d[0], 1), d[3], d[4]))
