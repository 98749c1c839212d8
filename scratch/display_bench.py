"""EXPERIMENT helper: time of the display chain of one 1080p frame (6 levels), fused (2 launches) vs the operator chain,
and of LineEndDisplayer.callback end to end."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysilent_b200 import LineEndDisplayer
rs = np.random.RandomState(5)
frame = rs.randint(0, 256, size=(1080, 1920, 3), dtype=np.uint8)
disp = LineEndDisplayer(zoom_ratio=2 ** .5)
res = disp.run_frames(torch.from_numpy(frame).cuda().unsqueeze(0), want_points=False)
gray = disp and None
from pysilent_b200.util.color import get_value_from_color
gray = get_value_from_color(res.padded_line_end)
for fused in (True, False):
    for _ in range(5): disp.display_tensors(res.orient, res.padded_line_end, gray, fused=fused)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(100): disp.display_tensors(res.orient, res.padded_line_end, gray, fused=fused)
    b.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("display chain fused=%s: device %.1f us, wall %.1f us per call" % (fused, a.elapsed_time(b) * 10, (t1 - t0) * 1e4))
for _ in range(5): disp.callback(frame)
t0 = time.perf_counter()
for _ in range(50): disp.callback(frame)
print("callback (1080p frame in, six tensors out): %.2f ms per frame" % ((time.perf_counter() - t0) * 20))
