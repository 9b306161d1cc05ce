import csv, subprocess, sys, collections, re
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = [line]; blocks.append(cur)
    elif cur is not None:
        cur.append(line)
b = blocks[which]
print(b[0][:120])
rd = list(csv.reader(b[1:]))
h = rd[0]; rows = [r for r in rd[1:] if len(r) == len(h)]
S = h.index("# Samples"); I = h.index("Instructions Executed"); SRC = h.index("Source")
tot_s = sum(float(r[S]) for r in rows); tot_i = sum(float(r[I]) for r in rows)
print("total samples", tot_s, "total warp-instr", tot_i, "static instr", len(rows))
ops = collections.Counter(); ops_s = collections.Counter()
for r in rows:
    op = r[SRC].strip().split()[0]
    if op.startswith("@"): op = r[SRC].strip().split()[1]
    op = op.split(".")[0]
    ops[op] += float(r[I]); ops_s[op] += float(r[S])
print("by opcode (share of executed instr | share of samples):")
for op, c in ops.most_common(22):
    print(f"  {op:12s} {100*c/tot_i:5.1f}%  {100*ops_s[op]/tot_s:5.1f}%")
print("hottest SASS by samples:")
for i, r in sorted(enumerate(rows), key=lambda t: -float(t[1][S]))[:top]:
    print(f"  {100*float(r[S])/tot_s:5.2f}%  exec={float(r[I]):>11.0f}  #{i:5d}  {r[SRC].strip()[:100]}")
