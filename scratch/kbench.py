"""Quick per-kernel device timing of the C3 step (EXPERIMENT helper, not the bench): prints pyramid / stack_a / stack_b /
emit ms for a batch, via the library's event hooks."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysilent_b200 import LineEndPipeline, _lib

B = int(os.environ.get("KB_BATCH", "64"))
shape = tuple(int(v) for v in os.environ.get("KB_SHAPE", "1080,1920").split(","))
scale = float(os.environ.get("KB_SCALE", str(2 ** .5)))
torch.cuda.set_device(0)
rs = np.random.RandomState(3)
frames = torch.from_numpy(rs.randint(0, 256, size=(B,) + shape + (3,), dtype=np.uint8)).cuda()
pipe = LineEndPipeline(zoom_ratio=scale)
plan = pipe.plan_for(frames)
n = B * plan.levels
bufs = (torch.empty((n, plan.h, plan.w, 3), dtype=torch.float32, device="cuda"),
        torch.empty((n, plan.h, plan.w, 3), dtype=torch.float32, device="cuda"),
        torch.empty((64 * n, 4), dtype=torch.int64, device="cuda"), torch.zeros(1, dtype=torch.int64, device="cuda"))
for _ in range(3):
    pipe.run_frames(frames, out=bufs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    pipe.run_frames(frames, out=bufs)
e1.record()
torch.cuda.synchronize()
total = e0.elapsed_time(e1) / 10
L = _lib.lib()
_lib.check(L.silent_plan_enable_timing(plan.handle, 1))
acc = np.zeros(5)
R = 6
for i in range(R):
    pipe.run_frames(frames, out=bufs)
    v = [ctypes.c_float() for _ in range(5)]
    _lib.check(L.silent_plan_stage_ms(plan.handle, ctypes.byref(v[0]), ctypes.byref(v[1]), ctypes.byref(v[2])))
    _lib.check(L.silent_plan_stack_split_ms(plan.handle, ctypes.byref(v[3]), ctypes.byref(v[4])))
    if i:
        acc += [x.value for x in v]
acc /= (R - 1)
print("%s step %.4f ms (%.0f fps) | pyramid %.4f stack_a %.4f stack_b %.4f emit %.4f" % (
    os.environ.get("KB_TAG", ""), total, B / total * 1e3, acc[0], acc[3], acc[4], acc[2]), flush=True)
