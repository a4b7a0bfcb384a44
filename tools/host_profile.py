import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import torch_motion_correction_b200 as tmc
dev = torch.device("cuda:0")
cfg = bench.WORKLOADS["c2"]
movie, _ = bench.synthetic_movie_gpu(cfg["t"], cfg["h"], cfg["w"], 1000, dev)
for _ in range(3):
    tmc.motion_correct(movie, 0.83, patch_sidelength=1024, n_iterations=100)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
    tmc.motion_correct(movie, 0.83, patch_sidelength=1024, n_iterations=100)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
