import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from pysilent_b200 import LineEndPipeline
B=64
pipe=LineEndPipeline(zoom_ratio=2**.5)
rs=np.random.RandomState(0)
hf=torch.from_numpy(rs.randint(0,256,size=(B,1080,1920,3),dtype=np.uint8)).pin_memory()
n=B*6
o=torch.empty((n,192,288,3),dtype=torch.float32).pin_memory().numpy()
l=torch.empty((n,192,288,3),dtype=torch.float32).pin_memory().numpy()
for _ in range(3): pipe.run_host(hf.numpy(),o,l)
torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(10): pipe.run_host(hf.numpy(),o,l)
torch.cuda.synchronize()
ms=(time.perf_counter()-t0)*100
print(os.environ.get("SILENT_CHUNK"), "e2e %.2f ms/step  %.0f fps"%(ms, B/ms*1e3), flush=True)
