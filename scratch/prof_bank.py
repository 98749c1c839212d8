"""Config C4 steps for an ncu capture of stack_bank_kernel: ncu ... -k regex:stack_bank -s 2 -c 1 python scratch/prof_bank.py"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysilent_b200 import LineEndPipeline
torch.cuda.set_device(0)
fr = torch.from_numpy(np.random.RandomState(4).randint(0, 256, (16, 2160, 3840, 3), dtype=np.uint8)).cuda()
pipe = LineEndPipeline(zoom_ratio=2 ** .5, orientations=8)
for _ in range(4):
    pipe.run_frames(fr)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    pipe.run_frames(fr)
b.record()
torch.cuda.synchronize()
print("C4: %.3f ms per 16 frames = %.0f frames/s" % (a.elapsed_time(b) / 10, 160e3 / a.elapsed_time(b)))
