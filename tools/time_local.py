import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import torch_motion_correction_b200 as tmc
dev = torch.device("cuda:0")
cfg = bench.WORKLOADS["c2"]
movie, _ = bench.synthetic_movie_gpu(cfg["t"], cfg["h"], cfg["w"], 1000, dev)
field = torch.zeros((2, 40, 6, 6), device=dev)
for n in (1, 20, 100, 200):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        tmc.estimate_local_motion(movie, 0.83, (1024, 1024), (3, 5, 5), field, n_iterations=n, grid_type="bspline")
        torch.cuda.synchronize(); print(n, rep, round((time.perf_counter() - t0) * 1e3, 2), "ms", flush=True)
