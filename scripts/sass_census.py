"""SASS opcode census of libaero_sm100.so per kernel (cuobjdump -sass): proves which kernels use tcgen05 / TMA.
usage: python scripts/sass_census.py > profiles/rNN_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "aero_gnn_b200", "libaero_sm100.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
ops = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "HMMA", "LDSM", "SYNCS", "ELECT"]
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name)
        counts.setdefault(cur, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        if op in ops:
            counts[cur][op] += 1
print("# SASS opcode census of aero_gnn_b200/libaero_sm100.so (cuobjdump -sass; sm_100a)")
print("# tcgen05.mma -> UTC*MMA, tcgen05.commit -> UTCBAR, tcgen05.ld/st -> LDTM/STTM, cp.async.bulk.tensor -> UTMALDG/UTMASTG,")
print("# tensormap prefetch -> UTMAPF, cp.async.bulk -> UBLKCP, mma.sync -> HMMA, ldmatrix -> LDSM, mbarrier -> SYNCS, elect.sync -> ELECT")
print("kernel," + ",".join(ops))
for k in sorted(counts):
    c = counts[k]
    if any(c[o] for o in ops):
        print(k + "," + ",".join(str(c[o]) for o in ops))
