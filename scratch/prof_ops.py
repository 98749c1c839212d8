"""The stand-alone operators of the reference's API on one 1080p frame (6 levels), for an ncu capture of the kernels that
are NOT on the bench path: pyramid_kernel (from_image), conv2d_kernel<K> (rgc / rgby / stripe / blur / end filters one by
one, 8-orientation bank), regulate / pad_inwards / value kernels, the stand-alone selection kernels, the display chain.
    ncu --set full --clock-control none -k regex:silent -c 60 -o gpurun_out/prof_ops python scratch/prof_ops.py"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysilent_b200 import LineEndPipeline, LineEndDisplayer
from pysilent_b200.util.zoom import from_image
torch.cuda.set_device(0)
frame = np.random.RandomState(6).randint(0, 256, (1080, 1920, 3), dtype=np.uint8)
reps = int(os.environ.get("PROF_REPS", "1"))
for _ in range(reps):
    pyr = from_image(frame, 3, (288, 192), 2 ** .5)   # pyramid_kernel
    pipe = LineEndPipeline(zoom_ratio=2 ** .5)
    res = pipe.run_unfused(pyr)                       # operator by operator: conv2d_kernel<3/7>, regulate, pad, value, selection
    bank = LineEndPipeline(zoom_ratio=2 ** .5, orientations=8).run_bank(pyr)
    disp = LineEndDisplayer(zoom_ratio=2 ** .5)
    for fused in (True, False):
        disp.display_tensors(res.orient, res.padded_line_end, res.gray, fused=fused)
    torch.cuda.synchronize()
print("done", tuple(res.orient.shape), tuple(bank.orient.shape))
