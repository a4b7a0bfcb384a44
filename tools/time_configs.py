"""Device-resident motion_correct timings of the BASELINE configs C2 / C3 / C4 (single GPU)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import torch_motion_correction_b200 as tmc
dev = torch.device("cuda:0")
CONFIGS = {
    "c2": dict(t=40, h=4096, w=4096, px=0.83, patch=1024, res=(3, 5, 5)),
    "c3": dict(t=60, h=4096, w=4096, px=0.83, patch=1024, res=(5, 6, 6)),
    "c4": dict(t=40, h=8192, w=8192, px=0.415, patch=1024, res=(3, 5, 5)),
}
for name in sys.argv[1:] or ["c2", "c3", "c4"]:
    c = CONFIGS[name]
    movie, _ = bench.synthetic_movie_gpu(c["t"], c["h"], c["w"], 7, dev)
    kw = dict(patch_sidelength=c["patch"], deformation_field_resolution=c["res"], n_iterations=100)
    out, field = tmc.motion_correct(movie, c["px"], **kw); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); out, field = tmc.motion_correct(movie, c["px"], **kw); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    print(name, c, "motion_correct", round(best, 2), "ms; field", tuple(field.shape), "finite", bool(torch.isfinite(out).all()),
          "max|field| A", round(float(field.abs().max()), 3), flush=True)
    del movie, out
    torch.cuda.empty_cache()
