"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list:
kernel, launches, total ms, share.  usage: launch_list_summary.py launches.csv [command line shown in the header]"""
import collections, csv, re, sys
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
rd = list(csv.reader(rows))
h = rd[0]
K, V, U = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rd[1:]:
    name = re.sub(r"\(.*", "", r[K])[:110]
    ns = float(r[V].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[U], 1.0)
    tot[name] += ns; cnt[name] += 1
total = sum(tot.values())
cmd = sys.argv[2] if len(sys.argv) > 2 else ""
print(f"# ncu launch list summary (gpu__time_duration.sum, --clock-control none): {cmd}")
print(f"# {sum(cnt.values())} launches captured, total {total / 1e6:.1f} ms (cold-cache, serialised: compare SHARES)")
print("kernel,launches,total_ms,share")
for name, ns in sorted(tot.items(), key=lambda kv: -kv[1])[:24]:
    print(f"{name},{cnt[name]},{ns / 1e6:.3f},{ns / total:.4f}")
