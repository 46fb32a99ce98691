import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import synthdata as synth
from dirb200 import nets
B, S = 64, 1024
net = nets.create_model("resnet101_rmac")
net.load_state_dict(synth.make_state_dict("resnet101_rmac", seed=0))
g = torch.Generator(device="cuda").manual_seed(1)
imgs = torch.randn((B, 3, S, S), generator=g, device="cuda")
ref = None
configs = [("flat", dict(stage_sched=0)),
           ("auto", dict(stage_sched=1)),
           ("A 1,1,2,5,16", dict(sub0=1, sub1=1, sub2=2, sub3=5, sub4=16)),
           ("B 2,2,4,9,32", dict(sub0=2, sub1=2, sub2=4, sub3=9, sub4=32)),
           ("C 2,4,8,16,64", dict(sub0=2, sub1=4, sub2=8, sub3=16, sub4=64)),
           ("D l3=9 only", dict(sub0=64, sub1=64, sub2=64, sub3=9, sub4=64)),
           ("E l3=5 only", dict(sub0=64, sub1=64, sub2=64, sub3=5, sub4=64)),
           ("F l3=16 only", dict(sub0=64, sub1=64, sub2=64, sub3=16, sub4=64)),
           ("G early 2,2,4", dict(sub0=2, sub1=2, sub2=4, sub3=64, sub4=64)),
           ("H early 4,4,8", dict(sub0=4, sub1=4, sub2=8, sub3=64, sub4=64)),
           ("I 8,8,16,32,64", dict(sub0=8, sub1=8, sub2=16, sub3=32, sub4=64)),
           ]
net.forward(imgs)
for name, opts in configs:
    for k in ("sub0", "sub1", "sub2", "sub3", "sub4"):
        net.set_backend_option_live(k, 0)
    net.set_backend_option_live("stage_sched", 1)
    for k, v in opts.items():
        net.set_backend_option_live(k, v)
    for _ in range(2):
        d = net.forward(imgs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        d = net.forward(imgs)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    if ref is None:
        ref = d.clone()
    same = bool(torch.equal(ref, d))
    print("%-18s %7.2f ms/step  %7.1f img/s  launches %5d  identical=%s" % (name, ms, B / ms * 1e3, net.last_launch_stats()[0], same), flush=True)
