"""Per-kernel device times of one motion_correct step on the benchmark movie (diagnostic; library timing facility):
python tools/time_stage_kernels.py [substring ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import torch_motion_correction_b200 as tmc
from torch_motion_correction_b200 import _lib

dev = torch.device("cuda:0")
size = int(os.environ.get("TMC_TIME_SIZE", "4096"))  # 8192: BASELINE config 4's frame size
movie, _ = bench.synthetic_movie_gpu(40, size, size, 1000, dev)
for _ in range(2):
    tmc.motion_correct(movie, 0.83, n_iterations=20)
torch.cuda.synchronize()
_lib.kernel_timing(True)
for _ in range(3):
    tmc.motion_correct(movie, 0.83, n_iterations=20)
torch.cuda.synchronize()
rep = _lib.kernel_timing_report()
_lib.kernel_timing(False)
keys = sys.argv[1:]
print({k: round(ms / 3, 4) for k, (n, ms) in rep.items() if not keys or any(s in k for s in keys)})
