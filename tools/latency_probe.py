import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import synthdata as synth
from dirb200 import nets, ops
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    ops.set_global_option(k, float(v))
    print("option", k, v)
net = nets.create_model("resnet101_rmac"); net.load_state_dict(synth.make_state_dict("resnet101_rmac", seed=0))
for (b, h, w) in ((1, 1024, 768), (1, 1024, 1024), (1, 512, 384), (4, 1024, 768), (8, 1024, 1024), (1, 224, 224), (16, 224, 224)):
    x = torch.randn((b, 3, h, w), device="cuda")
    for _ in range(3): net.forward(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 10
    for _ in range(n): net.forward(x)
    t_host = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / n
    print("B=%d %dx%d: %.2f ms/forward (host enqueue %.2f ms) -> %.1f img/s" % (b, h, w, t_all * 1e3, t_host * 1e3, b / t_all), flush=True)
