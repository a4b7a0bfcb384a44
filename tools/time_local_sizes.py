"""Per-iteration time of the spline optimiser against the number of (patch, tile) CTAs (wave quantisation probe)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import torch_motion_correction_b200 as tmc
dev = torch.device("cuda:0")
for size in (2048, 3072, 3584, 4096, 5120):
    movie, _ = bench.synthetic_movie_gpu(40, size, size, 1000, dev)
    f0, c = tmc.estimate_motion_cross_correlation_patches(movie, 0.83, patch_sidelength=1024)
    g = c.shape[1] * c.shape[2]
    res = {}
    for n in (100, 300):
        for rep in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            tmc.estimate_local_motion(movie, 0.83, (1024, 1024), (3, 5, 5), f0, n_iterations=n, grid_type="bspline")
            torch.cuda.synchronize(); res[n] = (time.perf_counter() - t0) * 1e3
    print(f"size {size} patches {g} CTAs {g * 26}: {(res[300] - res[100]) / 200 * 1e3:.1f} us per iteration", flush=True)
    del movie
    torch.cuda.empty_cache()
