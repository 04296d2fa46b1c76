"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import csv, sys, collections
path = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
iu = hdr.index("Metric Unit")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
    name = r[ik].split("(")[0].replace("<unnamed>::", "")
    tot[name] += v; cnt[name] += 1
total = sum(tot.values())
print(f"# per-kernel device time from {path} (ncu launch list; cold-cache, serialised: compare SHARES)")
print(f"{'kernel':60s} {'launches':>9s} {'total ms':>12s} {'share':>8s} {'avg ms':>10s}")
for k, v in tot.most_common():
    print(f"{k:60s} {cnt[k]:9d} {v:12.3f} {100*v/total:7.2f}% {v/cnt[k]:10.4f}")
print(f"{'TOTAL':60s} {sum(cnt.values()):9d} {total:12.3f}")
