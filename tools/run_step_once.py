"""One motion_correct step on the benchmark movie (target of ncu captures): python tools/run_step_once.py [iterations]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import torch_motion_correction_b200 as tmc

dev = torch.device("cuda:0")
movie, _ = bench.synthetic_movie_gpu(40, 4096, 4096, 1000, dev)
tmc.motion_correct(movie, 0.83, n_iterations=int(sys.argv[1]) if len(sys.argv) > 1 else 3)
torch.cuda.synchronize()
print("ok")
