"""How far the CUDA loss / gradient are from the fp64 oracle, and what that does to L-BFGS (debugging aid)."""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch_motion_correction_b200 as tmc
from oracle import reference_path as rp
from torch_motion_correction_b200.estimate_motion_optimizer import LocalMotionProblem
from torch_motion_correction_b200 import estimate_motion_optimizer as emo

dev = torch.device("cuda:0")
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "small.npz"))
movie = torch.as_tensor(g["movie"]); px = float(g["pixel_spacing"]); fr = tuple(float(v) for v in g["frequency_range"])
init = torch.as_tensor(g["xc_full_mean_except_current"])
gen = torch.Generator().manual_seed(3)
new = torch.randn((2, 3, 3, 3), generator=gen) * 0.3
batches = [[i] for i in range(16)]
for dt in (torch.float32, torch.float64):
    wl, wg = rp.loss_and_grad(movie.to(dt), px, (32, 32), (3, 3, 3), init.to(dt), new.to(dt), batches, frequency_range=fr, grid_type="bspline", loss_type="mse")
    print("oracle", dt, wl, float(wg.abs().max()))
    if dt == torch.float64: wl64, wg64 = wl, wg
    else: wl32, wg32 = wl, wg
print("oracle fp32 vs fp64: loss rel", abs(wl32 - wl64) / abs(wl64), "grad rel", float((wg32.double() - wg64).abs().max() / wg64.abs().max()))
for fused in (True, False):
    emo.FUSED_STEPS = fused
    prob = LocalMotionProblem(movie.to(dev), px, (32, 32), (3, 3, 3), init.to(dev), dev, 500, fr, "bspline", "mse")
    scale = torch.tensor([s / len(batches) for s in prob.patch_scales(batches)], dtype=torch.float32).to(dev)
    loss, grad = prob.loss_and_grad(new.to(dev), scale)
    wl, wg = wl64 / len(batches), wg64 / len(batches)
    print("cuda fused" if fused else "cuda generic", "loss rel", abs(float(loss) - wl) / abs(wl), "grad rel", float((grad.cpu().double() - wg).abs().max() / wg.abs().max()))
for fused in (True, False):
    emo.FUSED_STEPS = fused
    random.seed(1234)
    res, traj = tmc.estimate_local_motion(movie.to(dev), px, (32, 32), (3, 3, 3), init.to(dev), n_iterations=6, frequency_range=fr,
                                          return_trajectory=True, optimizer_type="lbfgs", grid_type="bspline", loss_type="mse")
    print("lbfgs fused" if fused else "lbfgs generic", "max diff px", float((res.cpu() - torch.as_tensor(g["local_lbfgs_bspline_mse"])).abs().max()) / px)
    print("  losses", [c.loss for c in traj.checkpoints]); print("  golden", list(g["local_lbfgs_bspline_mse_losses"]))
