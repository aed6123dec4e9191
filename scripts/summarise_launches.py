"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: total us, count, share."""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 14 and r[0].isdigit()]
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    name = re.sub(r"\(.*", "", r[4])
    tot[name] += float(r[14]) / 1e3
    cnt[name] += 1
s = sum(tot.values())
print("    total us  count   share  kernel")
for k, v in tot.most_common():
    print(f"{v:12.1f} {cnt[k]:6d} {100 * v / s:6.1f}%  {k[:110]}")
