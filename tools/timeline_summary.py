import csv, collections, sys
rows = list(csv.DictReader(open(sys.argv[1])))
dur = collections.OrderedDict(); prev = {}
for row in rows:
    st, k, t = row["stream"], row["kernel"], float(row["t_end_us"])
    if st in prev: dur.setdefault(k, []).append(t - prev[st])
    prev[st] = t
tot = 0
for k, v in dur.items():
    print(f"   {k:40s} n={len(v):3d} mean {sum(v)/len(v):8.1f} us  total/iter {sum(v)/6:8.1f}"); tot += sum(v)
print("total per iteration", tot / 6)
