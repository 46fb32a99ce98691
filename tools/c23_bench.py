"""A/B timing of the Bottleneck tail at the ResNet-101 64 x 1024^2 shapes: two kernels (3x3 halo + 1x1 with residual) vs
the fused kernel (one CTA per tile) vs the fused CTA-pair kernel (tcgen05 cta_group::2).  CUDA events, 10 repetitions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dirb200 import ops

DEV = "cuda:0"
for (cm, b, h, w) in ((256, 64, 64, 64), (128, 64, 128, 128), (64, 64, 256, 256)):
    r = np.random.RandomState(cm)
    t1 = torch.relu(torch.randn((b, h, w, cm), device=DEV)).half()
    res = torch.relu(torch.randn((b, h, w, 4 * cm), device=DEV)).half()
    w2 = torch.from_numpy((r.standard_normal((cm, cm, 3, 3)) * np.sqrt(2.0 / (9 * cm))).astype(np.float32))
    w3 = torch.from_numpy((r.standard_normal((4 * cm, cm, 1, 1)) * np.sqrt(2.0 / cm)).astype(np.float32))
    s2, h2 = torch.ones(cm, device=DEV), torch.zeros(cm, device=DEV)
    s3, h3 = torch.full((4 * cm,), 0.3, device=DEV), torch.zeros(4 * cm, device=DEV)
    w2p, w3p = ops.pack_conv_weight(w2).to(DEV), ops.pack_conv_weight(w3).to(DEV)

    def two():
        t2 = ops.conv_bn_act(t1, w2p, cm, 3, 3, 1, 1, s2, h2, None, True)
        return ops.conv_bn_act(t2, w3p, 4 * cm, 1, 1, 1, 0, s3, h3, res, True)

    variants = {"two kernels": two,
                "fused, pair": lambda: ops.conv_c23(t1, w2p, s2, h2, w3p, s3, h3, res, variant=1),
                "fused, single": lambda: ops.conv_c23(t1, w2p, s2, h2, w3p, s3, h3, res, variant=0)}
    ref = two()
    flops = 2.0 * b * h * w * (9.0 * cm * cm + 4.0 * cm * cm)
    for name, fn in variants.items():
        out = fn()
        torch.cuda.synchronize()
        same = bool(torch.equal(out, ref))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("Cm %3d %dx%d B=%d  %-14s %8.3f ms  %7.1f TFLOP/s  identical=%s" % (cm, h, w, b, name, ms, flops / (ms * 1e-3) / 1e12, same), flush=True)
