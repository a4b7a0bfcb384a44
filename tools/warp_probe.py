"""Probe of the TMA warp kernel on one small problem (debugging aid): python tools/warp_probe.py [h w t]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_motion_correction_b200 as tmc

h, w, t = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (300, 388, 5)
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(21)
img = torch.randn((t, h, w), generator=g).to(dev)
field = (torch.randn((2, 3, 4, 5), generator=g) * 2.0).to(dev)
total = tmc.correct_motion_sum(img, field, 1.1, grid_type="bspline")
torch.cuda.synchronize()
os.environ["TMC_WARP_TMA"] = "0"
want = tmc.correct_motion_sum(img, field, 1.1, grid_type="bspline")
torch.cuda.synchronize()
print("debug", os.environ.get("TMC_WARP_TMA_DEBUG"), "ok, rel diff", float((total - want).norm() / want.norm()))
