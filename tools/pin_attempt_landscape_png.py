"""Attempt to pin the oracle against the only numeric-ish artefact the reference ships for this path: the committed
render images/out/landscape_evolution.png of examples/landscape_evolution.rs (30 000 sites from
StdRng::from_seed([0;32]), relaxate_sites(1), k = 1, hull outlets, get_elevation per pixel, grey = (z / max * 255) as u8).

    python tools/pin_attempt_landscape_png.py [num_sites] [mean|none|centroid]      (this container only: reads /root/reference)

The whole chain is restated (site stream, voronoice's Lloyd step as the mean of the clipped cell's vertices -- from
memory, the crate is not vendored --, clipped Voronoi areas, builder graph, generate(), natural-neighbour render).
RESULT (round 1): texture, brightness range (max grey 249 vs 251) and iteration count look alike, but the drainage
pattern does not coincide (pixel correlation 0.3-0.4 = the shared centre-bright trend only) for any Lloyd variant.  The
pattern of iteration 1 is decided by the epsilon-noise on exactly these site positions, so one wrong link -- the crate
version that produced the committed image, voronoice's Lloyd/clip details, or one of the restated RNG streams -- is
enough, and the image cannot tell which.  Parity therefore stays UNPINNED.  Not used by tests.
"""
import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np
from scipy.spatial import Delaunay
from PIL import Image
from oracle import oracle as O
from tools import workloads as W

def clip_poly(poly, a, b, c):
    out=[]; k=len(poly)
    for i in range(k):
        p,q=poly[i],poly[(i+1)%k]
        fp=a*p[0]+b*p[1]-c; fq=a*q[0]+b*q[1]-c
        if fp<=0: out.append(p)
        if (fp<0<fq) or (fq<0<fp):
            t=fp/(fp-fq); out.append((p[0]+t*(q[0]-p[0]), p[1]+t*(q[1]-p[1])))
    return out

def voronoi_cells(pts, lo, hi):
    """per-site (mean of cell vertices, cell area) of the Voronoi diagram clipped to the box [lo,hi]."""
    n=pts.shape[0]
    dl=Delaunay(pts)
    tri,_=W._orient_ccw(pts, dl.simplices.astype(np.int64))
    a,b,c=pts[tri[:,0]],pts[tri[:,1]],pts[tri[:,2]]
    ex,ey=b[:,0]-a[:,0],b[:,1]-a[:,1]; fx,fy=c[:,0]-a[:,0],c[:,1]-a[:,1]
    d=2*(ex*fy-ey*fx); e2=ex*ex+ey*ey; f2=fx*fx+fy*fy
    cc=np.stack([a[:,0]+(fy*e2-ey*f2)/d, a[:,1]+(ex*f2-fx*e2)/d],1)
    inside=(cc[:,0]>=lo[0])&(cc[:,0]<=hi[0])&(cc[:,1]>=lo[1])&(cc[:,1]<=hi[1])
    # per-site: sum of circumcentres, count, any outside
    sx=np.zeros(n); sy=np.zeros(n); cnt=np.zeros(n); bad=np.zeros(n,bool)
    for k in range(3):
        v=tri[:,k]
        sx+=np.bincount(v,weights=cc[:,0],minlength=n); sy+=np.bincount(v,weights=cc[:,1],minlength=n)
        cnt+=np.bincount(v,minlength=n); bad|=np.bincount(v,weights=(~inside).astype(float),minlength=n)>0
    hull=np.unique(dl.convex_hull.reshape(-1)); bad[hull]=True
    mean=np.stack([sx/np.maximum(cnt,1), sy/np.maximum(cnt,1)],1)
    # areas of interior cells: sum over incident triangles of cross(cc_t - s, cc_next - s)/2 : use edge-based formula
    # each interior Delaunay edge (u,v) shared by triangles t1,t2: Voronoi edge cc1-cc2 contributes triangle (s, cc1, cc2) to both cells
    frm=tri.reshape(-1); to=tri[:,[1,2,0]].reshape(-1); tid=np.repeat(np.arange(tri.shape[0]),3)
    key=frm*n+to; rkey=to*n+frm
    order=np.argsort(key); pos=np.minimum(np.searchsorted(key[order],rkey),key.size-1)
    hit=key[order][pos]==rkey
    t2=np.where(hit, tid[order[pos]], -1)
    m=hit&(frm<to)
    u,v=frm[m],to[m]; p=cc[tid[m]]; q=cc[t2[m]]
    area=np.zeros(n)
    for s in (u,v):
        ps=pts[s]
        ar=0.5*np.abs((p[:,0]-ps[:,0])*(q[:,1]-ps[:,1])-(p[:,1]-ps[:,1])*(q[:,0]-ps[:,0]))
        area+=np.bincount(s,weights=ar,minlength=n)
    cent=np.zeros((n,2))
    for s_ in (u,v):
        ps=pts[s_]
        ar=0.5*np.abs((p[:,0]-ps[:,0])*(q[:,1]-ps[:,1])-(p[:,1]-ps[:,1])*(q[:,0]-ps[:,0]))
        cen=(ps+p+q)/3.0
        cent[:,0]+=np.bincount(s_,weights=ar*cen[:,0],minlength=n); cent[:,1]+=np.bincount(s_,weights=ar*cen[:,1],minlength=n)
    cent/=np.maximum(area,1e-300)[:,None]
    # clipped cells by half-plane clipping against Delaunay neighbours
    indptr,indices=dl.vertex_neighbor_vertices
    box=[(lo[0],lo[1]),(hi[0],lo[1]),(hi[0],hi[1]),(lo[0],hi[1])]
    for i in np.nonzero(bad)[0]:
        poly=box; s=pts[i]
        for j in indices[indptr[i]:indptr[i+1]]:
            t=pts[j]
            poly=clip_poly(poly, 2*(t[0]-s[0]), 2*(t[1]-s[1]), (t[0]**2+t[1]**2)-(s[0]**2+s[1]**2))
        P=np.array(poly)
        mean[i]=P.mean(0)
        cx=np.roll(x:=P[:,0],-1); cy=np.roll(y:=P[:,1],-1); cr=x*cy-cx*y; A6=3*cr.sum()
        cent[i]=((x+cx)*cr).sum()/A6, ((y+cy)*cr).sum()/A6
        x,y=P[:,0],P[:,1]
        area[i]=0.5*abs(np.dot(x,np.roll(y,-1))-np.dot(y,np.roll(x,-1)))
    return mean, area, bad, cent

num=int(sys.argv[1]) if len(sys.argv)>1 else 30000
mode=sys.argv[2] if len(sys.argv)>2 else 'mean'
lo,hi=(0.0,0.0),(100.0,100.0)
t0=time.time()
sites=O.random_sites(num, lo, hi, 0)
mean,_,bad,cent=voronoi_cells(sites,lo,hi)
print("lloyd", time.time()-t0, "clipped cells", bad.sum())
sites={'mean':mean,'none':sites,'centroid':cent}[mode]
_,area,bad,_c=voronoi_cells(sites,lo,hi)
print("total area", area.sum())
m=W.model_from_triangles(sites, Delaunay(sites).simplices)
m["areas"]=area
p=W.uniform_params(m["n"])
outlets=W.outlets_for(m,p)
initial=O.initial_elevations(p["base"])
t0=time.time()
e,it=O.generate(m,p["erodibility"],p["uplift"],None,outlets,initial)
print("generate", time.time()-t0, "iterations", it, "max", e.max())
W_=500
cols,rows=np.meshgrid(np.arange(W_,dtype=np.float64),np.arange(W_,dtype=np.float64))
q=np.stack([100.0*(cols.reshape(-1)/W_),100.0*(rows.reshape(-1)/W_)],1)
sites_,tri,he=W.triangulation_of(m)
t0=time.time()
z=O.nn_interpolate(sites_,tri,e,q,walk=True).reshape(W_,W_)
print("render", time.time()-t0)
img=np.zeros((W_,W_),np.uint8)
ok=~np.isnan(z)
img[ok]=np.floor(z[ok]/e.max()*255.0).astype(np.uint8)
ref=np.array(Image.open('/root/reference/images/out/landscape_evolution.png').convert('RGB'))[:,:,0]
print("ref shape", ref.shape, "ref max", ref.max(), "mine max", img.max())
np.save('/tmp/pin/mine_%d.npy'%num, img); np.save('/tmp/pin/z_%d.npy'%num, z)
d=img.astype(int)-ref.astype(int)
inner=(slice(50,450),slice(50,450))
print("corr all", np.corrcoef(img.reshape(-1),ref.reshape(-1))[0,1], "corr inner", np.corrcoef(img[inner].reshape(-1),ref[inner].reshape(-1))[0,1])
print("exact", (d==0).mean(), "within1", (np.abs(d)<=1).mean(), "within2", (np.abs(d)<=2).mean(), "inner within1", (np.abs(d[inner])<=1).mean())
Image.fromarray(img).save('/tmp/pin/mine_%d_%s.png'%(num,mode))
