"""Per-entry-point device time of motion_correct with and without a concurrent H2D copy."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import torch_motion_correction_b200 as tmc
from torch_motion_correction_b200 import _lib
dev = torch.device("cuda:0")
cfg = bench.WORKLOADS["c2"]
movie, _ = bench.synthetic_movie_gpu(cfg["t"], cfg["h"], cfg["w"], 1000, dev)
host = torch.empty(movie.shape, dtype=torch.float32, pin_memory=True); host.copy_(movie)
other = torch.empty_like(movie)
kw = dict(patch_sidelength=1024, deformation_field_resolution=(3, 5, 5), n_iterations=100)
tmc.motion_correct(movie, 0.83, **kw); torch.cuda.synchronize()
cs = torch.cuda.Stream()
for concurrent in (False, True):
    torch.cuda.synchronize()
    if concurrent:
        with torch.cuda.stream(cs):
            other.copy_(host, non_blocking=True)
    _lib.TIMING = {}
    t0 = time.perf_counter()
    tmc.motion_correct(movie, 0.83, **kw)
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    per = {k: round(sum(a.elapsed_time(b) for a, b in v), 2) for k, v in _lib.TIMING.items()}
    _lib.TIMING = None
    print("concurrent" if concurrent else "alone", "host-issue", round(t_issue * 1e3, 1), "ms", {k: v for k, v in sorted(per.items(), key=lambda kv: -kv[1])[:8]}, flush=True)
