import sys, os, json, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/model-predictive-control-for-bipedal-locomotion_b200')
import bench
from mpc_bipedal import engine
engine.LIBRARY_PATH = sys.argv[1]
from tools.riccati_probe import timed
dev=torch.device('cuda',0)
out={}
for B in (1024, 4096, 8192):
    cfg, z_max, z_min, F, tick = bench.sweep_inputs(B)
    zmin=torch.tensor(z_min,device=dev); zmax=torch.tensor(z_max,device=dev)
    plan=engine.LipmPlan(cfg.horizon,cfg.h,cfg.g,cfg.dt,cfg.Q,cfg.R)
    x0=torch.zeros(B,3,dtype=torch.float64,device=dev); y0=torch.zeros_like(x0)
    ft=torch.tensor(tick,device=dev)
    for name,Fs in (("push",F),("nopush",np.zeros_like(F))):
        imp=torch.tensor(cfg.dt*Fs/cfg.m,device=dev)
        r={}
        ms=timed(lambda: plan.rollout(1,x0,y0,zmin,zmax,impulse=imp,force_tick=ft,out=r,want_readout=True),reps=4)
        out[f"{name}{B}"]=round(ms,2)
    del plan
print(os.path.basename(sys.argv[1]), json.dumps(out), float(r["y_hist"].sum()))
