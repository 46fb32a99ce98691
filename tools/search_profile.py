import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dirb200 import ops
# (rows, queries, options): eps16 = -1 puts the filter threshold above every score (no candidate appends: isolates the
# cost of the epilogue's atomics), sample_rows trades seed-pass time for a tighter threshold (fewer candidates)
CASES = () if sys.argv[1:] == ["shard"] else ((125_000, 1000, {}),) if sys.argv[1:] == ["one"] else ((1_000_000, 1000, {}),) if sys.argv[1:] == ["c4"] else ((100_000, 70, {}),) if sys.argv[1:] == ["c3"] else (
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

if sys.argv[1:] in ([], ["shard"]):
    # One rank's compute of the 8-way sharded protocol (dist.ShardedIndex.search) on ONE GPU: search_begin with the
    # shard quota c = ceil(k / 8), [the MIN all-reduce is replaced by the local value], search_finish, merge of 8 lists.
    # What is left for N = 8 on top of these numbers: one 4 KB all-reduce and one 1.6 MB/rank all-gather.
    N, Q, K, G = 125_000, 1000, 100, 8
    g = torch.Generator(device="cuda").manual_seed(1)
    db, db16 = ops.l2_normalize(torch.randn((N, 2048), generator=g, device="cuda"), want_f16=True)
    q = ops.l2_normalize(torch.randn((Q, 2048), generator=g, device="cuda"))
    idx = ops.Index(db, db16=db16)
    idx.set_option("deferred_check", 1)
    idx.set_option("retries", 2)
    packed = torch.empty((2, Q, K), dtype=torch.int64, device="cuda")
    gathered = torch.empty((G, 2, Q, K), dtype=torch.int64, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    acc = [0.0, 0.0, 0.0]
    for it in range(8):
        ev[0].record()
        sel = idx.search_begin(q, K, (K + G - 1) // G)
        ev[1].record()
        idx.search_finish(q, K, sel, out=packed)
        ev[2].record()
        gathered[:] = packed                       # (stands in for the all-gather payload; not timed as a collective)
        ev[2].record()
        out = ops.topk_merge_packed(gathered, K)
        ev[3].record()
        torch.cuda.synchronize()
        if it >= 3:
            for j in range(3):
                acc[j] += ev[j].elapsed_time(ev[j + 1]) / 5
    idx.check()
    print("shard of 8: N %d Q %d k %d  begin %.3f ms  finish %.3f ms  merge %.3f ms  sum %.3f ms" % (N, Q, K, acc[0], acc[1], acc[2], sum(acc)),
          idx.stats(), flush=True)
