"""Aggregates an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line."""
import csv, sys
def f(x):
    try: return float(x)
    except: return 0.0
rows=[x for x in csv.reader(open(sys.argv[1]))]
top=int(sys.argv[2]) if len(sys.argv)>2 else 40
hdr=None; lines=[]
for x in rows:
    if x and x[0]=='Line No': 
        hdr=x; ix={}
        for i,h in enumerate(hdr): ix.setdefault(h,i)
        continue
    if hdr is None or not x or not x[0].isdigit(): continue
    lines.append((int(x[0]), x[1], f(x[ix['Instructions Executed']]), f(x[ix['# Samples']])))
ti=sum(l[2] for l in lines); ts=sum(l[3] for l in lines)
print('total inst %.0f samples %.0f'%(ti,ts))
agg={}
for l in lines:
    a=agg.setdefault(l[0],[l[1],0,0]); a[1]+=l[2]; a[2]+=l[3]
key=(lambda kv:-kv[1][2]) if len(sys.argv)<4 else (lambda kv:-kv[1][1])
for ln,(src,i,s) in sorted(agg.items(), key=key)[:top]:
    print('%4d  inst %5.1f%%  samp %5.1f%%  %s'%(ln,100*i/ti,100*s/ts,src.strip()[:120]))
