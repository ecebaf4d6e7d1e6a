"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): launches, total time and share per kernel."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, n = collections.Counter(), collections.Counter()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    k = r[ki].split("(")[0]
    tot[k] += v
    n[k] += 1
T = sum(tot.values())
print(f"# ncu launch list of `{sys.argv[2] if len(sys.argv) > 2 else '?'}` (gpu__time_duration.sum, --clock-control none)")
print("# per-launch times are serialised / cold-cache: compare SHARES, not absolutes")
print(f"# total device time of all launches: {T:.1f} ms")
print("kernel,launches,total_ms,share")
for k, v in tot.most_common():
    print(f"{k},{n[k]},{v:.3f},{v / T:.4f}")
