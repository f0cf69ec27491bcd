"""BASELINE configs[4] on one GPU: bench.py's param_sweep leg alone (131,072 scenarios, FP64 and reduced precision).
    python tools/param_sweep_probe.py [scenarios]"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from mpc_bipedal import engine  # noqa: E402

per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
torch.cuda.set_device(0)
engine.keep_pool_memory(0)
res = bench.param_sweep_leg(torch, None, engine, torch.device("cuda", 0), 0, 1, per_gpu=per_gpu)
print(json.dumps(res))
