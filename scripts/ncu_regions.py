"""Warp-stall samples and executed instructions of one kernel in an .ncu-rep, aggregated between synchronisation
landmarks (BAR / mbarrier SYNCS / tcgen05 commit) of the SASS: shows which phase of a persistent kernel the time goes to."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rd = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rd) if "# Samples" in r][0]
h = rd[hi]
rows = [r for r in rd[hi + 1:] if len(r) == len(h)]
S, SRC, I = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
ts = sum(float(r[S]) for r in rows)
ti = sum(float(r[I]) for r in rows)
acc = acci = 0.0
last = 0
print(f"total samples {ts:.0f}, warp instructions {ti:.0f}")
for i, r in enumerate(rows):
    acc += float(r[S]); acci += float(r[I])
    ops = [o for o in r[SRC].strip().split() if not o.startswith("@")]
    base = ops[0].split(".")[0] if ops else ""
    if base in ("BAR", "SYNCS", "EXIT", "UTCBAR") and acc / ts > 0.002:
        print(f"SASS [{last:4d}-{i:4d}] samples {100 * acc / ts:5.2f}%  instr {100 * acci / ti:5.2f}%  x{float(r[I]):.0f}  ends: {r[SRC].strip()[:56]}")
        acc = acci = 0.0
        last = i + 1
