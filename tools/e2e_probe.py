import sys, time, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/model-predictive-control-for-bipedal-locomotion_b200')
import numpy as np, torch
import bench
from mpc_bipedal.controllers import ZMPController
from mpc_bipedal import engine
B=4096
cfg, z_max, z_min, F, tick = bench.sweep_inputs(B)
ctrl = ZMPController(cfg)
hx0=np.zeros((B,3)); hy0=np.zeros((B,3))
for i in range(5):
    torch.cuda.synchronize(); t0=time.perf_counter()
    res = ctrl.generate_com_trajectory_batched(hx0, hy0, z_max, z_min, F_ext=F, force_tick=tick)
    torch.cuda.synchronize(); t1=time.perf_counter()
    print('full call', i, round((t1-t0)*1e3,1),'ms')
for i in range(3):
    torch.cuda.synchronize(); t0=time.perf_counter()
    res = ctrl.generate_com_trajectory_batched(hx0, hy0, z_max, z_min, F_ext=F, force_tick=tick, to_numpy=False)
    torch.cuda.synchronize(); t1=time.perf_counter()
    out = engine.to_host(res)
    t2=time.perf_counter()
    print('device part', round((t1-t0)*1e3,1),'ms; to_host', round((t2-t1)*1e3,1),'ms', sum(v.nbytes for v in out.values())/1e6,'MB')
