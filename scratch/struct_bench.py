import sys, os, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from conftest import structured_frame
from pysilent_b200 import LineEndPipeline
few = np.stack([structured_frame(900 + i, 1080, 1920) for i in range(8)])
frames = torch.from_numpy(np.concatenate([few] * 8, axis=0)).cuda()
pipe = LineEndPipeline(zoom_ratio=2 ** .5)
plan = pipe.plan_for(frames); plan.reserve(64); n = 64 * plan.levels
bufs = (torch.empty((n, plan.h, plan.w, 3), device="cuda"), torch.empty((n, plan.h, plan.w, 3), device="cuda"),
        torch.empty((64 * n, 4), dtype=torch.int64, device="cuda"), torch.zeros(1, dtype=torch.int64, device="cuda"))
for _ in range(3): pipe.run_frames(frames, out=bufs)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): pipe.run_frames(frames, out=bufs)
b.record(); torch.cuda.synchronize()
print("structured input: %.4f ms per 64-frame step" % (a.elapsed_time(b) / 10))
