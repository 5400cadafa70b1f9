import csv, subprocess, sys
rep, only = sys.argv[1], sys.argv[2]
groups = eval(sys.argv[3])  # list of (name, lo, hi)
out = subprocess.run(["ncu","-i",rep,"--page","source","--print-source","cuda,sass","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
ok=True; hdr=None; data={}
for r in rows:
    if r and r[0]=="Function Name": ok = only in r[1]; continue
    if not ok: continue
    if r and r[0]=="Line No": hdr=r; ie=hdr.index("Instructions Executed"); sm=hdr.index("# Samples"); continue
    if hdr is None or len(r)<=ie or not r[0]: continue
    try: ln=int(r[0]); n=int(r[ie]); s=int(r[sm] or 0)
    except ValueError: continue
    a=data.setdefault(ln,[0,0]); a[0]+=n; a[1]+=s
tot=sum(v[0] for v in data.values()); tots=sum(v[1] for v in data.values())
print("total inst",tot,"samples",tots)
acc={}
for ln,(n,s) in data.items():
    g="other"
    for name,lo,hi in groups:
        if lo<=ln<=hi: g=name;break
    a=acc.setdefault(g,[0,0]); a[0]+=n; a[1]+=s
for g,(n,s) in sorted(acc.items(), key=lambda kv:-kv[1][0]):
    print(f"{g:22s} {100*n/tot:5.1f}% inst {100*s/tots:5.1f}% samples")
oth=[(ln,v) for ln,v in data.items() if not any(lo<=ln<=hi for _,lo,hi in groups)]
for ln,(n,s) in sorted(oth,key=lambda kv:-kv[1][0])[:12]: print("   other L%d %.1f%% inst %.1f%% smp"%(ln,100*n/tot,100*s/tots))
