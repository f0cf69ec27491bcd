import sys, os, json, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/model-predictive-control-for-bipedal-locomotion_b200')
import bench
from mpc_bipedal import engine
engine.LIBRARY_PATH = sys.argv[1]
from tools.riccati_probe import timed
dev=torch.device('cuda',0)
B=4096
cfg, z_max, z_min, F, tick = bench.sweep_inputs(B)
zmin=torch.tensor(z_min,device=dev); zmax=torch.tensor(z_max,device=dev)
plan=engine.LipmPlan(cfg.horizon,cfg.h,cfg.g,cfg.dt,cfg.Q,cfg.R)
def run(sel, r):
    b=len(sel)
    x0=torch.zeros(b,3,dtype=torch.float64,device=dev); y0=torch.zeros_like(x0)
    imp=torch.tensor(cfg.dt*F[sel]/cfg.m,device=dev); ft=torch.tensor(tick[sel],device=dev)
    return lambda: plan.rollout(1,x0,y0,zmin,zmax,impulse=imp,force_tick=ft,out=r,want_readout=True)
r={}
out={"full": round(timed(run(np.arange(B), r), reps=4),2)}
it=r["iters"].cpu().numpy().reshape(B,-1).sum(axis=1)
order=np.argsort(-it)
for k in (1,64,1024):
    out[f"heaviest{k}"]=round(timed(run(np.sort(order[:k]),{}),reps=3),2)
out["B8192"]=None
print(os.path.basename(sys.argv[1]), json.dumps(out), float(r["y_hist"].sum()))
