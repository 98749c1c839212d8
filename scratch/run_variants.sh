python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python - <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from pysilent_b200 import LineEndPipeline
fr = torch.from_numpy(np.random.RandomState(4).randint(0,256,(16,2160,3840,3),dtype=np.uint8)).cuda()
pipe = LineEndPipeline(zoom_ratio=2**.5, orientations=8)
for _ in range(3): pipe.run_frames(fr)
torch.cuda.synchronize()
a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): pipe.run_frames(fr)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)/10
print("C4 fused: %.3f ms per 16 frames = %.0f fps" % (ms, 16e3/ms))
PY
