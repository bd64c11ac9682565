import json,sys
from collections import OrderedDict
d=json.load(open(sys.argv[1]))
L=d['launch_list']
g=OrderedDict()
for n,ms,f,b in L:
    k=(n,f,b)
    g.setdefault(k,[0,0.0])
    g[k][0]+=1; g[k][1]+=ms
tot=sum(v[1] for v in g.values())
print('total',tot)
for k,v in g.items():
    n,f,b=k
    if v[1] < float(sys.argv[2]) if len(sys.argv)>2 else 0: continue
    print(f"{n:18s} x{v[0]:3d} avg {v[1]/v[0]*1e3:8.1f} us  tot {v[1]:6.3f} ms  {f/ (v[1]/v[0]*1e-3)/1e12 if f else 0:7.1f} TF/s  {b/(v[1]/v[0]*1e-3)/1e9:7.1f} GB/s")
