import csv, subprocess, sys
rep=sys.argv[1]; kidx=sys.argv[2] if len(sys.argv)>2 else "0"
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass","--launch-skip",kidx,"--launch-count","1"],capture_output=True,text=True).stdout
fname,hdr,rows="",None,[]
cur=None
for r in csv.reader(out.splitlines()):
    if len(r)==2 and r[0] in ("File Name","File Path"): fname=r[1].split("/")[-1]
    elif r and r[0]=="Line No": hdr=r
    elif hdr and len(r)==len(hdr):
        if r[0]: cur=(fname,int(r[0]))
        else: rows.append((cur,r))
S,I,A=hdr.index("# Samples"),hdr.index("Instructions Executed"),hdr.index("Address")
# order SASS rows by address; attribute inlined helper lines to the enclosing position in the main kernel by address order
rows=[cr for cr in rows if cr[1][A].startswith("0x")]; rows.sort(key=lambda cr:int(cr[1][A],16))
ts=sum(float(r[S] or 0) for _,r in rows); ti=sum(float(r[I] or 0) for _,r in rows)
# walk in address order; track last seen line in main file
main=sys.argv[3]
bounds=[int(x) for x in sys.argv[4].split(",")]
names=sys.argv[5].split(",")
acc=[0.0]*len(names); acci=[0.0]*len(names)
last=0
for (f,l),r in rows:
    if f==main and l>=int(sys.argv[6] if len(sys.argv)>6 else 0): last=l
    k=0
    for i,b in enumerate(bounds):
        if last>=b: k=i
    acc[k]+=float(r[S] or 0); acci[k]+=float(r[I] or 0)
for n,a,b in zip(names,acc,acci): print(f"{n:28s} samples {100*a/ts:5.1f}%  instr {100*b/ti:5.1f}%")
