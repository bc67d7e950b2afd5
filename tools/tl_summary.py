"""Summarise a LNA_TIMELINE csv (stream, kernel, t_end_us): per-kernel span statistics.  usage: tl_summary.py file.csv [iters]"""
import collections, csv, sys
rows = list(csv.DictReader(open(sys.argv[1])))
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dur, prev = collections.OrderedDict(), {}
for r in rows:
    st, k, t = r["stream"], r["kernel"], float(r["t_end_us"])
    if st in prev:
        dur.setdefault(k, []).append(t - prev[st])
    prev[st] = t
tot = 0.0
for k, v in dur.items():
    print(f"   {k:34s} n={len(v):4d} mean {sum(v) / len(v):8.1f} us  sum/iter {sum(v) / iters:9.1f} us")
    tot += sum(v)
print(f"   all streams, per iteration {tot / iters:.1f} us")
