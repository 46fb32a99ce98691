import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dirb200 import ops
# (rows, queries, options): eps16 = -1 puts the filter threshold above every score (no candidate appends: isolates the
# cost of the epilogue's atomics), sample_rows trades seed-pass time for a tighter threshold (fewer candidates)
CASES = ((125_000, 1000, {}),) if sys.argv[1:] == ["one"] else ((1_000_000, 1000, {}),) if sys.argv[1:] == ["c4"] else ((100_000, 70, {}),) if sys.argv[1:] == ["c3"] else (
    (1_000_000, 1000, {}), (125_000, 1000, {}), (125_000, 1000, {"eps16": -1.0}), (125_000, 1000, {"sample_rows": 32768}),
    (1_000_000, 1000, {"eps16": -1.0}), (100_000, 70, {}))
for (N, Q, OPT) in CASES:
    g = torch.Generator(device="cuda").manual_seed(1)
    db, db16 = ops.l2_normalize(torch.randn((N, 2048), generator=g, device="cuda"), want_f16=True)
    q = ops.l2_normalize(torch.randn((Q, 2048), generator=g, device="cuda"))
    idx = ops.Index(db, db16=db16)
    idx.set_option("profile", 1)
    for k_, v_ in OPT.items():
        idx.set_option(k_, v_)
    for _ in range(3):
        idx.search(q, 100)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        idx.search(q, 100)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print("N", N, "Q", Q, OPT, "wall ms/search %.3f" % (dt * 1e3), {k: round(v, 3) for k, v in idx.profile().items()}, idx.stats(), flush=True)
    del idx, db, db16
    torch.cuda.empty_cache()
