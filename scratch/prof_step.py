"""Four C3 steps (3 warm-up + 1) for an ncu capture: ncu ... -k regex:silent -s 18 -c 6 python scratch/prof_step.py"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysilent_b200 import LineEndPipeline
B = int(os.environ.get("KB_BATCH", "64"))
torch.cuda.set_device(0)
rs = np.random.RandomState(3)
frames = torch.from_numpy(rs.randint(0, 256, size=(B, 1080, 1920, 3), dtype=np.uint8)).cuda()
pipe = LineEndPipeline(zoom_ratio=2 ** .5)
plan = pipe.plan_for(frames)
n = B * plan.levels
bufs = (torch.empty((n, plan.h, plan.w, 3), dtype=torch.float32, device="cuda"),
        torch.empty((n, plan.h, plan.w, 3), dtype=torch.float32, device="cuda"),
        torch.empty((64 * n, 4), dtype=torch.int64, device="cuda"), torch.zeros(1, dtype=torch.int64, device="cuda"))
torch.cuda.synchronize()
for _ in range(int(os.environ.get("PROF_STEPS", "4"))):
    pipe.run_frames(frames, out=bufs)
torch.cuda.synchronize()
print("done")
