import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import differt2d_b200 as d
from differt2d_b200 import functional as F
from oracle import ref_torch as R
from tests import helpers as H
from tests.test_gpu_parity import _vertex_scene, _grid, _cfg
method, mode = sys.argv[1], sys.argv[2]
sc=_vertex_scene(); osc=H.oracle_scene_from_product(sc)
X,Y=_grid(sc,14,16,"jitter"); grid=np.stack([X,Y],-1).reshape(-1,2)
xys,kinds,phis=sc.packed_objects(); fixed=np.stack([p.xy for p in sc.transmitters.values()])
C=50; x0=np.random.default_rng(1234).random((C,2),dtype=np.float32)
cands=R.all_path_candidates(7,0,2)
Ztot=np.zeros(len(grid)); 
bad=[]
for ci,c in enumerate(cands):
    k=len(c)
    # isolate candidate via order + filter
    filt=tuple(j for j in range(7) if j not in set(c.tolist()))
    sub=R.all_path_candidates(7,k,k,filter_nodes=filt)
    # x0 rows for the sub-list
    x0s=np.stack([x0[[i for i,cc in enumerate(cands) if cc.tolist()==s.tolist()][0]] for s in sub]) if k>0 else x0[:1]
    cfg=_cfg(mode,min_order=k,max_order=k,method=method,steps=100,filter_nodes=filt)
    x0k=x0s[:, :max(k,1)].copy() if k>0 else None
    Z,v=F.power_fwd(cfg,xys,fixed,grid,kinds=kinds,phis=phis,x0=x0k,alpha=100.0,want_valid=True,device="cuda")
    _,vo,fo=R.valid_masks(osc,osc.transmitters["tx"],torch.from_numpy(grid),method=method,min_order=k,max_order=k,filter_nodes=filt,x0=x0k,steps=100,approx=mode!="hard",alpha=100.0)
    v=v.cpu().numpy()[0]; vo=vo.float().numpy()
    si=[i for i,s in enumerate(sub) if s.tolist()==c.tolist()][0]
    dv=np.abs(v[:,si]-vo[:,si]); 
    if dv.max()>1e-3: bad.append((c.tolist(), int((dv>1e-3).sum()), float(dv.max())))
print("candidates with validity diffs:", bad)
