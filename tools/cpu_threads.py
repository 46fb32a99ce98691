import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import synthdata as synth
from oracle import dir_oracle as O
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "torch threads default", torch.get_num_threads())
sd = synth.make_state_dict("resnet101_rmac", seed=0)
x = synth.make_images(1, 1024, 1024, seed=1, smooth=False)
for nt in (8, 16, 32, 64):
    torch.set_num_threads(nt)
    O.extract(x[:, :, :256, :256], sd, "resnet101_rmac")
    t0 = time.perf_counter(); O.extract(x, sd, "resnet101_rmac"); dt = time.perf_counter() - t0
    print("threads", nt, "sec/img", round(dt, 3), flush=True)
