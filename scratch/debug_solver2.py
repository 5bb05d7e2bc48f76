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
cfg=_cfg(mode,max_order=2,method=method,steps=100,grid_cols=16)
Z,v=F.power_fwd(cfg,xys,fixed,grid,kinds=kinds,phis=phis,x0=x0,alpha=100.0,want_valid=True,device="cuda")
_,vo,fo=R.valid_masks(osc,osc.transmitters["tx"],torch.from_numpy(grid),method=method,max_order=2,x0=x0,steps=100,approx=mode!="hard",alpha=100.0)
v=v.cpu().numpy()[0]; vo=vo.float().numpy(); fo=fo.numpy()
Zo=(vo*fo).sum(-1); Zg=Z.cpu().numpy()[0]
rel=np.abs(Zg-Zo)/np.maximum(np.abs(Zo),1e-3)
print("rel err quantiles", np.quantile(rel,[0.5,0.9,0.95,0.99,1.0]))
print("flag mismatch", np.mean(v!=vo), "per cand:", {tuple(cands[i].tolist()): int((v[:,i]!=vo[:,i]).sum()) for i in range(C) if (v[:,i]!=vo[:,i]).any()})
worst=np.argsort(-rel)[:5]
for r in worst:
    print("rx",grid[r],"Zg",Zg[r],"Zo",Zo[r],"valid cands gpu",[cands[i].tolist() for i in np.nonzero(v[r])[0]],"oracle",[ (cands[i].tolist(), float(fo[r,i])) for i in np.nonzero(vo[r])[0]])
