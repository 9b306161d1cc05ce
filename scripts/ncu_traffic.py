"""Write profiles/rNN_ncu_traffic.json: DRAM bytes per launch of the fused block kernels from `ncu --set full`
reports, stamped with the commit and the sha256 of the kernel sources they were captured at (bench.py refuses a
capture whose sources differ from the tree: roofline.traffic = null).
usage: python scripts/ncu_traffic.py out.json N E dtype kind=report.ncu-rep:launch_index [kind=...]"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

out, N, E, dtype = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
res = {"workload": {"N": N, "E": E, "dtype": dtype},
       "commit": subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip(),
       "sources": bench.TRAFFIC_SOURCES, "sources_sha256": bench.sources_sha256(),
       "how": "ncu --set full --clock-control none, dram__bytes_read.sum / dram__bytes_write.sum of one launch"}
for spec in sys.argv[5:]:
    kind, rest = spec.split("=")
    rep, idx = rest.rsplit(":", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2 + int(idx)]
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))

    def bytes_of(key):
        v = float(d[key].replace(",", ""))
        return int(v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[key]])
    res[kind] = {"read_bytes": bytes_of("dram__bytes_read.sum"), "write_bytes": bytes_of("dram__bytes_write.sum"),
                 "duration_ms_under_ncu": float(d["gpu__time_duration.sum"].replace(",", "")) *
                 {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[u["gpu__time_duration.sum"]], "kernel": d["Kernel Name"][:80]}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
