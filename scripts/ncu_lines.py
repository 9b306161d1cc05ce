"""Hottest CUDA source lines of the first kernel in an .ncu-rep (needs -lineinfo and --import-source on):
stall samples and executed warp instructions per source line, with the dominant stall reasons.
usage: python scripts/ncu_lines.py report.ncu-rep [top=30] [kernel index=0]"""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
kidx = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", kidx,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
fname, hdr, lines = "", None, []
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0]:
        lines.append((fname, r))
S, I = hdr.index("# Samples"), hdr.index("Instructions Executed")
st = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ts = sum(float(r[S] or 0) for _, r in lines) or 1.0
ti = sum(float(r[I] or 0) for _, r in lines) or 1.0
print(f"total samples {ts:.0f}, warp instructions {ti:.0f}")
for f, r in sorted(lines, key=lambda fr: -float(fr[1][S] or 0))[:top]:
    reasons = sorted(((float(r[i] or 0), hdr[i][6:]) for i in st), reverse=True)[:3]
    rs = " ".join(f"{n}:{100 * v / max(float(r[S] or 1), 1):.0f}%" for v, n in reasons if v > 0)
    print(f"{100 * float(r[S] or 0) / ts:5.1f}% smp {100 * float(r[I] or 0) / ti:5.1f}% ins  {f}:{r[0]:>4s}  {r[1].strip()[:90]}  [{rs}]")
