"""How many grad_value reduction rows (corner pixels) of the bench workload are DISTINCT when the samples of Q
neighbouring queries of one head are pooled -- the ceiling of any scheme that merges rows before they leave the SM.
    python tools/count_shared_rows.py [grid|init|random]
Printed ratios are distinct rows / all rows: per (query, head, level) over its points, and over Q = 2 ... 128 consecutive
queries for the same sample slot and for the same level."""
import torch, numpy as np, sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dist = sys.argv[1] if len(sys.argv)>1 else 'grid'
value, loc, attn, go = bench.make_inputs(torch, 1, 0, dist)
shapes = bench.COCO_SHAPES
lsi,S = bench.level_start(shapes)
loc = loc[0].numpy()   # [Lq, M, L, P, 2]
Lq,M,L,P,_ = loc.shape
pix = np.full((Lq,M,L,P,4), -1, dtype=np.int64)
for l,(H,W) in enumerate(shapes):
    x = loc[:,:,l,:,0]*W-0.5; y = loc[:,:,l,:,1]*H-0.5
    x0 = np.floor(x).astype(np.int64); y0=np.floor(y).astype(np.int64)
    k=0
    for dy in (0,1):
        for dx in (0,1):
            xx=x0+dx; yy=y0+dy
            ok=(xx>=0)&(xx<W)&(yy>=0)&(yy<H)&(x>-1)&(y>-1)&(x<W)&(y<H)
            p = lsi[l]+yy*W+xx
            pix[:,:,l,:,k]=np.where(ok,p,-1); k+=1
total = (pix>=0).sum()
print(dist,'rows total', total, 'per sample', total/(Lq*M*L*P))
def distinct(groups):  # groups: array [..., n] of pixel ids; count distinct >=0 per leading index
    g = np.sort(groups, axis=-1)
    d = (g[...,1:]!=g[...,:-1]) & (g[...,1:]>=0)
    first = g[...,0]>=0
    return d.sum()+first.sum()
# (0) per-pair per-level (what in-group dedup could reach at best: all points of a level)
print('pair-level all-points merge', distinct(pix.reshape(Lq,M,L,P*4))/total)
# pair all samples
print('pair all samples', distinct(pix.reshape(Lq,M,L*P*4))/total)
for Q in (2,4,8,32,128):
    n=(Lq//Q)*Q
    p = pix[:n].reshape(n//Q,Q,M,L,P,4)
    # same slot across Q queries
    a = p.transpose(0,2,3,4,1,5).reshape(n//Q,M,L,P,Q*4)
    # same level across Q queries all points
    b = p.transpose(0,2,3,1,4,5).reshape(n//Q,M,L,Q*P*4)
    print('Q',Q,'same-slot',round(distinct(a)/total,3),'same-level',round(distinct(b)/total,3))
