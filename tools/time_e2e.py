import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import torch_motion_correction_b200 as tmc
dev = torch.device("cuda:0")
cfg = bench.WORKLOADS["c2"]
movie, _ = bench.synthetic_movie_gpu(cfg["t"], cfg["h"], cfg["w"], 1000, dev)
host = torch.empty(movie.shape, dtype=torch.float32, pin_memory=True); host.copy_(movie)
out = torch.empty((cfg["h"], cfg["w"]), dtype=torch.float32, pin_memory=True)
del movie
kw = dict(patch_sidelength=1024, deformation_field_resolution=(3, 5, 5), n_iterations=100)
def run(n):
    for _ in tmc.motion_correct_many((host for _ in range(n)), 0.83, device=dev, out_host=out, **kw): pass
run(2); torch.cuda.synchronize()
for n in (1, 2, 5, 10):
    t0 = time.perf_counter(); run(n); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    print(n, "movies:", round(dt, 1), "ms total", round(dt / n, 1), "ms/movie", flush=True)
