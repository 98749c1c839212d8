"""Small end-to-end case for compute-sanitizer --tool memcheck (one tool per gpurun call)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysilent_b200 import LineEndPipeline, LineEndDisplayer
rs = np.random.RandomState(0)
for shape, center, scale, batch in (((240, 320), (96, 64), 1.5, 3), ((200, 301 - 1), (52, 36), 1.3, 2), ((480, 640), (288, 192), 1.3, 1)):
    frames = rs.randint(0, 256, size=(batch,) + shape + (3,), dtype=np.uint8)
    pipe = LineEndPipeline(output_size=center, zoom_ratio=scale)
    res = pipe.run_frames(torch.from_numpy(frames).cuda())
    host = pipe.run_host(frames)
    assert np.array_equal(host.orient, res.orient.cpu().numpy(), equal_nan=True)
    disp = LineEndDisplayer(output_size=center, zoom_ratio=scale)
    out = disp.callback(frames[0])
    print(shape, center, "levels", res.orient.shape[0] // batch, "points", len(res.points), "display", len(out))
bank = LineEndPipeline(output_size=(96, 64), zoom_ratio=1.5, orientations=8)
r8 = bank.run_frames(torch.from_numpy(rs.randint(0, 256, size=(1, 240, 320, 3), dtype=np.uint8)).cuda())
torch.cuda.synchronize()
print("bank", tuple(r8.orient.shape), "ok")
