import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import synthdata as synth
from dirb200 import nets
net = nets.create_model("resnet101_rmac"); net.load_state_dict(synth.make_state_dict("resnet101_rmac", seed=0))
x = torch.randn((64, 3, 1024, 1024), device="cuda")
net.forward(x)
ref = None
for name, opts in (("base", {}), ("res_variant=1 <256,2,6>", {"res_variant": 1}), ("res_variant=2 <128,4,6>", {"res_variant": 2}), ("base again", {"res_variant": 0}), ("pdl=0", {"pdl": 0})):
    for k, v in opts.items():
        net.set_backend_option_live(k, v)
    for _ in range(2): d = net.forward(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4): d = net.forward(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 4
    if ref is None: ref = d.clone()
    print("%-26s %.2f ms/step %.1f img/s identical=%s" % (name, ms, 64 / ms * 1e3, bool(torch.equal(ref, d))), flush=True)
