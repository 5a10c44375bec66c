"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel.  usage: launches.py csv [out.csv]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hi]; ci = {n: i for i, n in enumerate(h)}
agg = collections.defaultdict(lambda: [0, 0.0]); order = []
for r in rows[hi + 1:]:
    if len(r) < len(h) or r[ci['Metric Name']] != 'gpu__time_duration.sum': continue
    name = r[ci['Kernel Name']].split('(')[0].replace('void ', '').replace('scnerf::', '')
    v = float(r[ci['Metric Value']]); u = r[ci['Metric Unit']]
    v = v / 1e3 if u in ('ns', 'nsecond') else (v if u in ('us', 'usecond') else v * 1e3)
    agg[name][0] += 1; agg[name][1] += v; order.append((name, v))
tot = sum(v[1] for v in agg.values())
lines = [f"# total {tot / 1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches (cold-cache, serialised: compare SHARES)",
         "kernel,launches,total_ms,share,avg_ms"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{k},{v[0]},{v[1] / 1e3:.3f},{v[1] / tot:.4f},{v[1] / v[0] / 1e3:.4f}")
print("\n".join(lines[:16]))
if len(sys.argv) > 2:
    open(sys.argv[2], 'w').write("\n".join(lines) + "\n")
# per-launch sequence of the big kernels in the last step
big = [(n, v) for n, v in order if v > 300]
print("last big launches (ms):", [(n.split('::')[-1][:22], round(v / 1e3, 2)) for n, v in big[-8:]])
