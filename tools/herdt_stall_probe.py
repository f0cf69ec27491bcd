import sys, os, numpy as np, torch, json
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/model-predictive-control-for-bipedal-locomotion_b200')
from mpc_bipedal.config import MPCConfig
from mpc_bipedal.controllers import ZMPController
from mpc_bipedal.generators.speed_generation import SpeedTrajectoryGenerator
cfg=MPCConfig(); cfg.method="herdt"; cfg.speed_generation="classic"
vx,vy,states=SpeedTrajectoryGenerator(cfg).generate_speed_and_state(save_footsteps=False)
v=np.stack([vx,vy],1); st=np.array(states)
for name,opts in (("default",{}),("nostall",{"stall_window":100000,"max_iter":20000,"accept_eps":-1.0}),("noaccept",{"accept_eps":-1.0})):
    ctrl=ZMPController(cfg); ctrl.herdt_options=opts
    batch=ctrl.prepare_herdt(st)
    n=len(st)
    out=batch.rollout(np.zeros((2,3)),np.zeros((2,3)),v,n_ticks=n-1)
    # per-tick: rerun tick loop to get per tick iters: use internal loop again but record
    print(name,'fallback ticks',out['status'].tolist(),'accepted',out['accepted'].tolist(),'iters',out['iters'].tolist(),'final com',out['x_hist'][0,-1,0].item(),out['y_hist'][0,-1,0].item())
# per-tick detail with nostall
ctrl=ZMPController(cfg); ctrl.herdt_options={"stall_window":100000,"max_iter":20000,"accept_eps":-1.0}
batch=ctrl.prepare_herdt(st)
lay=batch.layout; dev=batch.plan.device
import mpc_bipedal.herdt as H
x=torch.zeros((1,3),dtype=torch.float64,device=dev); y=x.clone()
fc=torch.zeros((1,2),dtype=torch.float64,device=dev); fc[:,1]=cfg.foot_spread; air=fc.clone()
N=cfg.horizon
vref=torch.as_tensor(np.concatenate([v,np.repeat(v[-1:],N,0)])).to(dev)
rows=[]
for i in range(len(st)-1):
    side=int(lay.foot_side[i]); has=lay.m[i]>0
    o=batch.timeline.tick(i,x,y,fc,vref,side,batch.polygons[side] if has else None,air=air.contiguous(),warm=0 if i==0 else (2 if lay.new_run[i] else 1))
    x,y=o['next_x'],o['next_y']
    if has: air=air+(1.0/batch.nb_steps[i][0])*(o['f1']-air)
    if lay.commit[i]:
        fc=o['f1'].clone() if has else air.clone(); air=fc.clone()
    rows.append((i,int(lay.current[i]),int(lay.m[i]),o['status'].tolist(),o['iters'].tolist()))
for r in rows[236:]: print(r)
