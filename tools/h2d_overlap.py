"""H2D copy rate of one movie while motion_correct runs on another (does compute slow the DMA down?)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import torch_motion_correction_b200 as tmc
dev = torch.device("cuda:0")
cfg = bench.WORKLOADS["c2"]
movie, _ = bench.synthetic_movie_gpu(cfg["t"], cfg["h"], cfg["w"], 1000, dev)
host = torch.empty(movie.shape, dtype=torch.float32, pin_memory=True); host.copy_(movie)
other = torch.empty_like(movie)
kw = dict(patch_sidelength=1024, deformation_field_resolution=(3, 5, 5), n_iterations=100)
tmc.motion_correct(movie, 0.83, **kw); torch.cuda.synchronize()
cs = torch.cuda.Stream()
def copy_ms(concurrent):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    with torch.cuda.stream(cs):
        s.record(cs); other.copy_(host, non_blocking=True); e.record(cs)
    if concurrent:
        for _ in range(2): tmc.motion_correct(movie, 0.83, **kw)
    torch.cuda.synchronize()
    return s.elapsed_time(e)
for c in (False, True, False, True):
    print("concurrent compute" if c else "copy alone        ", round(copy_ms(c), 2), "ms", flush=True)
# compute time while a copy is running
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
with torch.cuda.stream(cs): other.copy_(host, non_blocking=True)
s.record(); tmc.motion_correct(movie, 0.83, **kw); e.record(); torch.cuda.synchronize()
print("compute during copy", round(s.elapsed_time(e), 2), "ms")
