import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
agg=[]; fn=None
def I(x):
    try: return int(x)
    except: return 0
for r in rows:
    if not r: continue
    if r[0]=='File Path': fn=r[1].split('/')[-1]; continue
    if r[0] in('Function Name','Line No'): continue
    if r[0]!='' and r[0].isdigit():
        agg.append((fn,int(r[0]),r[1].strip(),I(r[7]),I(r[4])))
tot=sum(a[3] for a in agg); ts=sum(a[4] for a in agg)
print('total inst',tot,'samples',ts)
n=int(sys.argv[2]) if len(sys.argv)>2 else 45
key=(lambda a:-a[3]) if len(sys.argv)<4 else (lambda a:-a[4])
for a in sorted(agg,key=key)[:n]:
    print(f"{a[0]:14s}:{a[1]:4d} {100*a[3]/tot:5.1f}% inst {100*a[4]/ts:5.1f}% stall  {a[2][:110]}")
