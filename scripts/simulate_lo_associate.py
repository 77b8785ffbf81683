"""CPU simulation of lo_associate's column-grid search (columns / points visited per query and pass) on a bench scan pair;
the numbers quoted in DESIGN.md section 10 item 2.  usage: python scripts/simulate_lo_associate.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pyoracle as O
from vloam_b200 import synth
s=synth.ScanStream(1234,n_cols=2048)
sr0=O.scan_registration(s.scan(0)); sr1=O.scan_registration(s.scan(1))
def sim(name,Q,T,is_corner):
    Tp=T[:,:3].astype(np.float64); ring=T[:,3].astype(int); n=len(Tp)
    minx,miny=Tp[:,0].min(),Tp[:,1].min()
    ix=np.floor(Tp[:,0]-minx).astype(int); iy=np.floor(Tp[:,1]-miny).astype(int)
    nx,ny=ix.max()+1,iy.max()+1
    cells={}
    for j,(a,b) in enumerate(zip(ix,iy)): cells.setdefault((a,b),[]).append(j)
    stats=dict(cols1=[],pts1=[],shells1=[],cols2=[],pts2=[],shells2=[])
    for q in Q[:,:3].astype(np.float64):
        qx,qy=int(np.floor(q[0]-minx)),int(np.floor(q[1]-miny))
        def lb(cx,cy):
            x0,y0=minx+cx,miny+cy
            dx=max(x0-q[0],q[0]-(x0+1),0); dy=max(y0-q[1],q[1]-(y0+1),0)
            return dx*dx+dy*dy
        def saferad(k):
            return min(q[0]-(minx+qx-k),(minx+qx+k+1)-q[0],q[1]-(miny+qy-k),(miny+qy+k+1)-q[1])
        # phase 1
        best=np.inf;bj=-1;cols=0;pts=0
        order=sorted([(0 if (a,b)==(0,0) else lb(qx+a,qy+b)+1e-12,a,b) for a in(-1,0,1) for b in(-1,0,1)])
        for l,a,b in order:
            if (a,b)!=(0,0) and l>best: break
            c=cells.get((qx+a,qy+b),[])
            if (a,b)!=(0,0) and not c: continue
            cols+=1;pts+=len(c)
            if c:
                d=((Tp[c]-q)**2).sum(1); m=d.argmin()
                if d[m]<best: best=d[m];bj=c[m]
        k=1;sh=0
        while True:
            R=saferad(k)
            if R>=5 or (np.isfinite(best) and R>0 and best<=R*R): break
            k+=1;sh+=1
            for a in range(-k,k+1):
                for b in range(-k,k+1):
                    if max(abs(a),abs(b))!=k: continue
                    c=cells.get((qx+a,qy+b),[])
                    cols+=1;pts+=len(c)
                    if c:
                        d=((Tp[c]-q)**2).sum(1); m=d.argmin()
                        if d[m]<best: best=d[m];bj=c[m]
        stats['cols1'].append(cols);stats['pts1'].append(pts);stats['shells1'].append(sh)
        if not (best<25): continue
        idc=ring[bj]
        # phase 2: window rings idc-2..idc+2 (approx), classes
        def classes(c):
            c=[j for j in c if abs(ring[j]-idc)<=2 and j!=bj]
            return c
        k2=k3=np.inf;cols=0;pts=0
        def consider(c):
            nonlocal k2,k3,pts
            c=classes(c);pts+=len(c)
            for j in c:
                d=((Tp[j]-q)**2).sum()
                if d>=25: continue
                same=(ring[j]<=idc) if j>bj else (ring[j]>=idc)
                if is_corner:
                    if not same: k2=min(k2,d)
                else:
                    if same: k2=min(k2,d)
                    else: k3=min(k3,d)
        lim=lambda: max(k2 if np.isfinite(k2) else 25, (k3 if np.isfinite(k3) else 25) if not is_corner else 0)
        for l,a,b in order:
            if (a,b)!=(0,0) and l>lim(): break
            c=cells.get((qx+a,qy+b),[])
            if (a,b)!=(0,0) and not c: continue
            cols+=1;consider(c)
        k=1;sh=0
        while True:
            R=saferad(k)
            if R>=5: break
            if R>0 and k2<=R*R and (is_corner or k3<=R*R): break
            k+=1;sh+=1
            for a in range(-k,k+1):
                for b in range(-k,k+1):
                    if max(abs(a),abs(b))!=k: continue
                    cols+=1;consider(cells.get((qx+a,qy+b),[]))
        stats['cols2'].append(cols);stats['pts2'].append(pts);stats['shells2'].append(sh)
    print(name,"queries",len(Q))
    for k,v in stats.items():
        v=np.array(v); print("  %-8s mean %.1f median %.0f p90 %.0f max %d  total %d"%(k,v.mean(),np.median(v),np.percentile(v,90),v.max(),v.sum()))
    sh=np.array(stats['shells2']); c2=np.array(stats['cols2'])
    for kk in range(0,5):
        sel=sh==kk
        if sel.any(): print("    phase2 extra shells=%d: n=%d, share of phase-2 columns %.2f"%(kk,sel.sum(),c2[sel].sum()/c2.sum()))
sim("corner",sr1.cornerPointsSharp,sr0.cornerPointsLessSharp,True)
sim("plane",sr1.surfPointsFlat,sr0.surfPointsLessFlat,False)
