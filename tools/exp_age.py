"""How the step time of the bench workload develops with the age of the games (diagnostic):
ms/step of agarcl_batch_step over windows of 100 env-steps, plus live cells / pellets of a few instances."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from agarcl_b200 import make_cfg
from agarcl_b200.batch import Batch

N = 4096
total = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
cfg = make_cfg(n_instances=N, device=0, **bench.WORKLOAD)
b = Batch(cfg)
b.seed(np.arange(N, dtype=np.uint64) + np.uint64(1))
b.reset()
gen = torch.Generator(device="cuda"); gen.manual_seed(1234)
dxdy = (torch.rand((16, N, 2), device="cuda", generator=gen) * 2 - 1).float().contiguous()
act = torch.randint(0, 3, (16, N), device="cuda", generator=gen, dtype=torch.int32).contiguous()
s = torch.cuda.current_stream().cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for w in range(total // 100):
    e0.record()
    for i in range(100):
        b.set_actions_device(dxdy[i % 16].data_ptr(), act[i % 16].data_ptr(), s)
        b.step(s)
    e1.record(); torch.cuda.synchronize()
    cells = []; pel = []; flags = 0
    for k in (0, 1000, 2000, 4095):
        sv = b.download_state(k)
        cells.append(int(sv.players["n_cells"].sum())); pel.append(int(sv.hdr["n_pellets"])); flags |= int(sv.hdr["flags"])
    print(f"steps {w*100:5d}..{w*100+99:5d}  {e0.elapsed_time(e1)/100:.4f} ms/step  cells {cells} pellets {pel} flags {flags}", flush=True)
hist = np.zeros(40, int); multi = []
for k in range(0, N, 32):
    nc = b.download_state(k).players["n_cells"]
    for v in nc: hist[min(int(v), 39)] += 1
    multi.append(int((nc >= 2).sum()))
print("n_cells histogram over 128 instances:", {i: int(h) for i, h in enumerate(hist) if h})
print("multi-cell players per instance: mean", np.mean(multi), "max", max(multi))
b.close()
