import numpy as np, sys
a=np.array([list(map(int,l.split())) for l in open(sys.argv[1])])
for cta in [0,60,120]:
    r=a[cta]
    items=[]
    for it in range(0,100):
        s=r[8+it*8:8+it*8+8]
        if s[0]<0 and s[2]<0: break
        items.append(s)
    items=np.array(items)
    d=np.diff(items[:,2])
    print("cta",cta,"n",len(items),"cycles/item",d[4:].mean(), "issue span",(items[4:,2]-items[4:,1]).mean(), "wait ops",(items[4:,1]-items[4:,0]).mean(),"epilogue",(items[4:,4]-items[4:,3]).mean(), "total cycles", items[-1,4]-items[0,0])
