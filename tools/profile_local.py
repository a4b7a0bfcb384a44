#!/usr/bin/env python
"""A few eager optimiser iterations of the C2 workload between cudaProfilerStart/Stop (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import torch_motion_correction_b200 as tmc
dev = torch.device("cuda:0")
cfg = bench.WORKLOADS["c2"]
movie, _ = bench.synthetic_movie_gpu(cfg["t"], cfg["h"], cfg["w"], 1000, dev)
field = torch.zeros((2, 40, 6, 6), device=dev)
tmc.estimate_local_motion(movie, 0.83, (1024, 1024), (3, 5, 5), field, n_iterations=3, grid_type="bspline", return_trajectory=True)
torch.cuda.synchronize(); torch.cuda.profiler.start()
tmc.estimate_local_motion(movie, 0.83, (1024, 1024), (3, 5, 5), field, n_iterations=3, grid_type="bspline", return_trajectory=True)
torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("done")
