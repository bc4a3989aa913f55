"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of GPU time)."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
tot, cnt = collections.Counter(), collections.Counter()
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
    name = row["Kernel Name"].replace("mtg::<unnamed>::", "").replace("void ", "")
    name = name.split("(")[0]
    tot[name] += v
    cnt[name] += 1
s = sum(tot.values())
print(f"{'kernel':60s} {'launches':>8s} {'total_us':>10s} {'share':>7s} {'avg_us':>8s}")
for n, v in tot.most_common(40):
    print(f"{n[:60]:60s} {cnt[n]:8d} {v:10.1f} {100*v/s:6.1f}% {v/cnt[n]:8.2f}")
print(f"{'TOTAL':60s} {sum(cnt.values()):8d} {s:10.1f}")
