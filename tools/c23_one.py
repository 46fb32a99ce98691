import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dirb200 import ops
DEV = "cuda:0"
cm, b, h, w = 256, 64, 64, 64
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 1
r = np.random.RandomState(cm)
t1 = torch.relu(torch.randn((b, h, w, cm), device=DEV)).half()
res = torch.relu(torch.randn((b, h, w, 4 * cm), device=DEV)).half()
w2 = torch.from_numpy((r.standard_normal((cm, cm, 3, 3)) * np.sqrt(2.0 / (9 * cm))).astype(np.float32))
w3 = torch.from_numpy((r.standard_normal((4 * cm, cm, 1, 1)) * np.sqrt(2.0 / cm)).astype(np.float32))
s2, h2 = torch.ones(cm, device=DEV), torch.zeros(cm, device=DEV)
s3, h3 = torch.full((4 * cm,), 0.3, device=DEV), torch.zeros(4 * cm, device=DEV)
w2p, w3p = ops.pack_conv_weight(w2).to(DEV), ops.pack_conv_weight(w3).to(DEV)
for _ in range(4):
    ops.conv_c23(t1, w2p, s2, h2, w3p, s3, h3, res, variant=variant)
torch.cuda.synchronize()
print("done")
