"""Phase timing of the end-to-end push sweep (NumPy in, NumPy out): where do the slow calls lose their time?"""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import bench
from mpc_bipedal import engine
B = 4096
cfg, z_max, z_min, F, tick = bench.sweep_inputs(B)
plan = engine.get_plan(cfg.horizon, cfg.h, cfg.g, cfg.dt, cfg.Q, cfg.R)
hx0 = np.zeros((B, 3)); impulse = cfg.dt * F / cfg.m
pool = engine.HostPool()
cache = None
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
import subprocess, threading
rows = []
proc = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=timestamp,clocks.sm,power.draw,pstate", "--format=csv,noheader,nounits", "-lms", "20"],
                        stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append((time.perf_counter(), l.strip())) for l in proc.stdout], daemon=True).start()
marks = []
variant = sys.argv[1] if len(sys.argv) > 1 else "numpy_in+to_host"
print("variant:", variant)
if variant.startswith("device_in"):
    d = lambda a, dt: torch.as_tensor(a, dtype=dt).cuda()
    hx0, z_min, z_max, impulse, tick = d(hx0, torch.float64), d(z_min, torch.float64), d(z_max, torch.float64), d(impulse, torch.float64), d(tick, torch.int32)
for i in range(14):
    if i == 7:
        print("-- now with 256 MiB of L2-flush writes between calls (as the device-timed loop does)")
    if i >= 7:
        flush.fill_(1)
    t0 = time.perf_counter()
    out = plan.rollout(engine.MODE_STRICT, hx0, hx0, z_min, z_max, impulse=impulse, force_tick=tick, out=cache)
    cache = out
    t1 = time.perf_counter()
    if "spin" in variant:
        torch.cuda.synchronize()
    elif "blocking" in variant:
        ev = torch.cuda.Event(blocking=True); ev.record(); ev.synchronize()
    t2 = time.perf_counter()
    if variant.endswith("to_host"):
        res = engine.to_host({k: v for k, v in out.items() if not k.startswith("_")}, pool)
    t3 = time.perf_counter()
    marks.append((t1, t2))
    print(f"call {i:2d}: enqueue {1e3*(t1-t0):7.1f} ms | kernel wait {1e3*(t2-t1):7.1f} ms | to_host {1e3*(t3-t2):6.1f} ms")

proc.terminate()
for k, (a, b) in enumerate(marks):
    clk = [r[1].split(",")[1].strip() + "/" + r[1].split(",")[3].strip() for r in rows if a <= r[0] <= b]
    print("call", k, "sm clock/pstate samples during the kernel:", clk[:3], "...", clk[-3:], "n =", len(clk))
