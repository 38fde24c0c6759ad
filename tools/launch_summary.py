"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): python tools/launch_summary.py file.csv [last_n]"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr, L = None, []
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    try:
        L.append((r[hdr.index("Kernel Name")], r[hdr.index("Grid Size")], float(r[-1].replace(",", "")) / 1e3))
    except (ValueError, IndexError):
        pass
if len(sys.argv) > 2:
    L = L[-int(sys.argv[2]):]
tot = sum(v for _, _, v in L)
agg, cnt = collections.Counter(), collections.Counter()
for n, g, v in L:
    n = n.replace("mgb::", "").replace("<unnamed>::", "").replace("unnamed>::", "")[:70]
    agg[n] += v
    cnt[n] += 1
print("%-72s %8s %12s %7s" % ("kernel", "launches", "total_us", "share"))
for k, v in agg.most_common():
    print("%-72s %8d %12.1f %6.1f%%" % (k, cnt[k], v, 100 * v / tot))
print("%-72s %8d %12.1f" % ("total", len(L), tot))
