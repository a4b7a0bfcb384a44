"""A few iterations of the spline optimiser on the benchmark movie (target of ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import torch_motion_correction_b200 as tmc

dev = torch.device("cuda:0")
movie, _ = bench.synthetic_movie_gpu(40, 4096, 4096, 1000, dev)
f0, c = tmc.estimate_motion_cross_correlation_patches(movie, 0.83, patch_sidelength=1024)
tmc.estimate_local_motion(movie, 0.83, (1024, 1024), (3, 5, 5), f0, n_iterations=int(sys.argv[1]) if len(sys.argv) > 1 else 6, grid_type="bspline")
torch.cuda.synchronize()
print("ok")
