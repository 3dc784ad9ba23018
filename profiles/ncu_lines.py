import csv, subprocess, sys, collections
rep, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
mix = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows = list(csv.reader(mix.splitlines()))
hdr=None; tot=0; out=[]
for r in rows:
    if r and r[0]=='Line No': hdr=r; ie=hdr.index('Instructions Executed'); ss=hdr.index('# Samples'); continue
    if hdr and r and r[0].isdigit() and len(r)>ie:
        ln=int(r[0]); s=int(r[ss] or 0); i=int(r[ie] or 0); tot+=s
        if lo<=ln<=hi and (s or i): out.append((ln,s,i))
src=open('/root/repo/chapterhouseqe_b200/csrc/device_code.cuh').read().split('\n')
agg=collections.Counter(); aggi=collections.Counter()
for ln,s,i in out: agg[ln]+=s; aggi[ln]+=i
print("total samples",tot, "in range", sum(agg.values()))
for ln in sorted(agg, key=lambda k:-agg[k])[:int(sys.argv[4]) if len(sys.argv)>4 else 40]:
    print(f"{ln:5d} {agg[ln]:7d} smp {aggi[ln]:9d} inst  {src[ln-1].strip()[:110]}")
