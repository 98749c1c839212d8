"""Throughput of BASELINE config C4 (3840x2160, 8 levels, 8 orientations, batch 16) through the per-operator kernels."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysilent_b200 import LineEndPipeline
B = int(os.environ.get("KB_BATCH", "16"))
frames = torch.from_numpy(np.random.RandomState(4).randint(0, 256, size=(B, 2160, 3840, 3), dtype=np.uint8)).cuda()
pipe = LineEndPipeline(zoom_ratio=2 ** .5, orientations=8)
for _ in range(2):
    res = pipe.run_frames(frames)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    res = pipe.run_frames(frames)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("C4: %d frames, %d levels, %.3f ms/step, %.0f frames/s, points %d" % (B, res.orient.shape[0] // B, ms, B / ms * 1e3, len(res.points)))
# per-operator timing
from pysilent_b200 import _ops, _lib
from pysilent_b200.util.zoom.from_image import image_to_zoom_tensor
f = pipe.bank_filters()
def t(fn, *a, **k):
    torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record(); out = fn(*a, **k); e.record(); torch.cuda.synchronize(); return out, s.elapsed_time(e)
pyr, t0 = t(image_to_zoom_tensor, frames, 3, (288, 192), 2 ** .5)
a, t1 = t(_ops.conv2d, pyr, f["rgc"], post=_lib.POST_RELU)
b, t2 = t(_ops.conv2d, a, f["rgby"], post=_lib.POST_RELU)
c, t3 = t(_ops.conv2d, b, f["stripe"], post=_lib.POST_RELU)
d, t4 = t(_ops.regulate, c, f["blur"], 1.0, .1)
e, t5 = t(_ops.conv2d, d, f["end"], post=_lib.POST_RELU_CLIP, clip_max=255.0)
print("pyramid %.3f rgc %.3f rgby %.3f stripe(3->8) %.3f regulate(8) %.3f end(8->8) %.3f ms" % (t0, t1, t2, t3, t4, t5))
