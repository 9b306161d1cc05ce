import csv, subprocess, sys
rep = sys.argv[1]; which = int(sys.argv[2]); step = int(sys.argv[3]) if len(sys.argv) > 3 else 250
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = [line]; blocks.append(cur)
    elif cur is not None:
        cur.append(line)
rd = list(csv.reader(blocks[which][1:])); h = rd[0]; rows = [r for r in rd[1:] if len(r) == len(h)]
S = h.index("# Samples"); I = h.index("Instructions Executed"); SRC = h.index("Source")
ts = sum(float(r[S]) for r in rows); ti = sum(float(r[I]) for r in rows)
for b in range(0, len(rows), step):
    seg = rows[b:b + step]
    s = sum(float(r[S]) for r in seg); i = sum(float(r[I]) for r in seg)
    marks = [r[SRC].strip().split()[0 if not r[SRC].strip().startswith('@') else 1] for r in seg]
    special = sorted(set(m.split('.')[0] for m in marks if m.split('.')[0] in ("UTCHMMA", "LDTM", "SYNCS", "BAR", "LDG", "STG", "UBLKCP", "UTCBAR", "FENCE", "MUFU", "SHFL", "RED", "ATOM")))
    print(f"[{b:5d}] samples {100*s/ts:5.1f}%  instr {100*i/ti:5.1f}%  {' '.join(special)}")
