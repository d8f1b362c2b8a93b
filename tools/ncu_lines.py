"""Per-source-line instruction and stall-sample shares from an `ncu --page source --csv` export (lines above a threshold)."""
import csv, sys
def f(x):
    try: return float(x)
    except: return 0.0
rows=[x for x in csv.reader(open(sys.argv[1]))]
thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.004
hdr=None; agg={}
for x in rows:
    if x and x[0] in ('Line No','#'):
        hdr=x; ix={}
        for i,h in enumerate(hdr): ix.setdefault(h,i)
        continue
    if hdr is None or not x or not x[0].isdigit(): continue
    a=agg.setdefault(int(x[0]),[x[1],0,0]); a[1]+=f(x[ix['Instructions Executed']]); a[2]+=f(x[ix['# Samples']])
ti=sum(a[1] for a in agg.values()); ts=sum(a[2] for a in agg.values())
print('total inst %.0f samples %.0f'%(ti,ts))
cum=0
for ln in sorted(agg):
    s,i,sm=agg[ln]
    cum+=i
    if i/ti>thr or sm/ts>thr:
        print('%4d inst %5.2f%% cum %5.1f%% samp %5.2f%% %s'%(ln,100*i/ti,100*cum/ti,100*sm/ts,s.strip()[:100]))
