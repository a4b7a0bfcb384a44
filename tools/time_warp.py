import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import torch_motion_correction_b200 as tmc
dev = torch.device("cuda:0")
cfg = bench.WORKLOADS["c2"]
movie, _ = bench.synthetic_movie_gpu(cfg["t"], cfg["h"], cfg["w"], 1000, dev)
g = torch.Generator().manual_seed(0)
field = (torch.randn((2, 3, 5, 5), generator=g) * 3.0).to(dev)
for rep in range(4):
    torch.cuda.synchronize(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); out = tmc.correct_motion_sum(movie, field, 0.83, grid_type="bspline"); e.record(); torch.cuda.synchronize()
    print("correct_motion_sum", round(s.elapsed_time(e), 3), "ms", flush=True)
for rep in range(2):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); out = tmc.correct_motion(movie, field, 0.83, grid_type="bspline"); e.record(); torch.cuda.synchronize()
    print("correct_motion (stack)", round(s.elapsed_time(e), 3), "ms", flush=True)
