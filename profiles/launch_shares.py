"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total, share, average.
usage: python profiles/launch_shares.py gpurun_out/launches.csv > profiles/rNN_launch_shares.txt"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
tot = collections.OrderedDict()
for r in rows[1:]:
    try:
        us = float(r[ix["Metric Value"]].replace(",", "")) / (1000.0 if r[ix["Metric Unit"]] == "ns" else 1.0)
    except ValueError:
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])[:72]
    t = tot.setdefault(name, [0, 0.0]); t[0] += 1; t[1] += us
total = sum(v[1] for v in tot.values())
print("# per-launch times under ncu are cold-cache / serialised: compare SHARES, not absolute times")
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:72s} launches {n:4d} total {us:10.1f} us  share {100 * us / total:5.1f}%  avg {us / n:8.1f} us")
