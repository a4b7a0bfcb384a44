import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import torch_motion_correction_b200 as tmc
from torch_motion_correction_b200 import _lib
dev = torch.device("cuda:0")
cfg = bench.WORKLOADS["c2"]
movie, _ = bench.synthetic_movie_gpu(cfg["t"], cfg["h"], cfg["w"], 1000, dev)
px = 0.83
def timed(name, fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        _lib.TIMING = {}
        t0 = time.perf_counter(); out = fn(); t_host = time.perf_counter() - t0
        torch.cuda.synchronize(); t_all = time.perf_counter() - t0
        dev_ms = sum(a.elapsed_time(b) for v in _lib.TIMING.values() for a, b in v); _lib.TIMING = None
        best = min(best, t_all)
        print(f"{name:28s} host-issue {t_host*1e3:7.2f} ms   wall {t_all*1e3:7.2f} ms   device-sum {dev_ms:7.2f} ms")
    return out
g = timed("estimate_global_motion", lambda: tmc.estimate_global_motion(movie, px))
f, _ = timed("xc_patches(with global)", lambda: tmc.estimate_motion_cross_correlation_patches(movie, px, patch_sidelength=1024, deformation_field=g.clone()))
l = timed("estimate_local_motion(100)", lambda: tmc.estimate_local_motion(movie, px, (1024, 1024), (3, 5, 5), f, n_iterations=100, grid_type="bspline"))
timed("correct_motion_sum", lambda: tmc.correct_motion_sum(movie, l, px, grid_type="bspline"))
timed("motion_correct(100 it)", lambda: tmc.motion_correct(movie, px, patch_sidelength=1024, n_iterations=100))
