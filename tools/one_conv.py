"""One convolution shape in a loop (for ncu captures): python tools/one_conv.py Cin Cout k res [B H W] [key=value ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dirb200 import ops

a = [x for x in sys.argv[1:] if "=" not in x]
for kv in [x for x in sys.argv[1:] if "=" in x]:
    k, v = kv.split("=")
    ops.set_global_option(k, float(v))
cin, cout, k, use_res = int(a[0]), int(a[1]), int(a[2]), int(a[3])
b, h, w = (int(a[4]), int(a[5]), int(a[6])) if len(a) >= 7 else (64, 64, 64)
r = np.random.RandomState(0)
x = torch.relu(torch.randn((b, h, w, cin), device="cuda")).half()
res = torch.relu(torch.randn((b, h, w, cout), device="cuda")).half() if use_res else None
wt = torch.from_numpy((r.standard_normal((cout, cin, k, k)) * np.sqrt(2.0 / (cin * k * k))).astype(np.float32))
wp = ops.pack_conv_weight(wt).to("cuda")
s, sh = torch.full((cout,), 0.5, device="cuda"), torch.zeros(cout, device="cuda")
for _ in range(3):
    y = ops.conv_bn_act(x, wp, cout, k, k, 1, k // 2, s, sh, res, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    y = ops.conv_bn_act(x, wp, cout, k, k, 1, k // 2, s, sh, res, True)
e1.record()
torch.cuda.synchronize()
print("conv %d->%d k%d res=%d @%dx%dx%d: %.1f us/launch" % (cin, cout, k, use_res, b, h, w, e0.elapsed_time(e1) / 20 * 1e3))
