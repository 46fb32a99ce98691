import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dirb200 import ops
CASES = ((125_000, 1000),) if sys.argv[1:] == ["one"] else ((1_000_000, 1000), (500_000, 1000), (250_000, 1000), (125_000, 1000), (125_000, 1024), (100_000, 70))
for (N, Q) in CASES:
    g = torch.Generator(device="cuda").manual_seed(1)
    db, db16 = ops.l2_normalize(torch.randn((N, 2048), generator=g, device="cuda"), want_f16=True)
    q = ops.l2_normalize(torch.randn((Q, 2048), generator=g, device="cuda"))
    idx = ops.Index(db, db16=db16)
    idx.set_option("profile", 1)
    for _ in range(3):
        idx.search(q, 100)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        idx.search(q, 100)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print("N", N, "Q", Q, "wall ms/search %.3f" % (dt * 1e3), {k: round(v, 3) for k, v in idx.profile().items()}, idx.stats(), flush=True)
    del idx, db, db16
    torch.cuda.empty_cache()
