import sys, os, json, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/model-predictive-control-for-bipedal-locomotion_b200')
import bench
from mpc_bipedal import engine
engine.LIBRARY_PATH = sys.argv[1]
from tools.riccati_probe import timed
dev=torch.device('cuda',0)
for B in (512, 1024, 2048, 8192):
    cfg, z_max, z_min, F, tick = bench.sweep_inputs(B)
    zmin=torch.tensor(z_min,device=dev); zmax=torch.tensor(z_max,device=dev)
    plan=engine.LipmPlan(cfg.horizon,cfg.h,cfg.g,cfg.dt,cfg.Q,cfg.R)
    x0=torch.zeros(B,3,dtype=torch.float64,device=dev); y0=torch.zeros_like(x0)
    imp=torch.tensor(cfg.dt*F/cfg.m,device=dev); ft=torch.tensor(tick,device=dev)
    res=[]
    for chunks in (1,4):
        plan.set_walk_chunks(chunks)
        for lanes in (0,1,2,4,8):
            plan.set_riccati_lanes(lanes)
            r={}
            ms=timed(lambda: plan.rollout(1,x0,y0,zmin,zmax,impulse=imp,force_tick=ft,out=r,want_readout=True),reps=3)
            res.append((chunks,lanes,round(ms,2)))
    print(B, res, flush=True)
