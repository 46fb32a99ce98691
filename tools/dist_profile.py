#!/usr/bin/env python
"""Phase timing of the sharded search under torchrun (one rank per GPU): 1M x 2048 rows split over the ranks, 1000 queries,
k = 100.  CUDA events around every phase of dist.ShardedIndex.search (max over ranks), then the sustained rate of
back-to-back searches with the status check deferred."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
import torch.distributed as dist

from dirb200 import ops
from dirb200.dist import ShardedIndex, shard_rows, shard_quota, all_gather_packed

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, Q, K = int(os.environ.get("DP_ROWS", "1000000")), 1000, 100
s0, s1 = shard_rows(N, world, rank)
g = torch.Generator(device="cuda").manual_seed(100 + rank)
db, db16 = ops.l2_normalize(torch.randn((s1 - s0, 2048), generator=g, device="cuda"), want_f16=True)
gq = torch.Generator(device="cuda").manual_seed(7)
q = ops.l2_normalize(torch.randn((Q, 2048), generator=gq, device="cuda"))
index = ShardedIndex(db, row_offset=s0, db16_local=db16)
for _ in range(3):
    index.search(q, K)
torch.cuda.synchronize()
names = ["begin", "allreduce_min", "finish", "allgather", "merge"]
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
acc = [0.0] * 5
index.local.set_option("deferred_check", 1)
k_shard = shard_quota(K, index.shard_sizes)
REP = 10
for it in range(REP):
    dist.barrier()
    ev[0].record()
    sel = index.local.search_begin(q, K, k_shard)
    ev[1].record()
    if world > 1:
        dist.all_reduce(sel, op=dist.ReduceOp.MIN)
    ev[2].record()
    packed = torch.empty((2, Q, K), dtype=torch.int64, device="cuda")
    index.local.search_finish(q, K, sel, out=packed)
    ev[3].record()
    gathered = all_gather_packed(packed)
    ev[4].record()
    out = ops.topk_merge_packed(gathered, K) if world > 1 else None
    ev[5].record()
    torch.cuda.synchronize()
    for j in range(5):
        acc[j] += ev[j].elapsed_time(ev[j + 1]) / REP
index.local.check()
t = torch.tensor(acc + [sum(acc)], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
# sustained: back-to-back searches, one synchronisation at the end
dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
STEPS = 30
e0.record()
for _ in range(STEPS):
    index.search(q, K, check=False)
e1.record()
torch.cuda.synchronize()
index.check()
ms = torch.tensor([e0.elapsed_time(e1) / STEPS], device="cuda", dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
# the same through the peer-memory exchange (one C call per search, no NCCL)
index.enable_peer_exchange(max_q=Q, max_k=K)
ref_s, ref_i = index.search(q, K)
index.xchg, keep = None, index.xchg
nccl_s, nccl_i = index.search(q, K)
index.xchg = keep
same = bool(torch.equal(ref_i, nccl_i) and torch.equal(ref_s, nccl_s))
for _ in range(3):
    index.search(q, K, check=False)
dist.barrier()
torch.cuda.synchronize()
e0.record()
for _ in range(STEPS):
    index.search(q, K, check=False)
e1.record()
torch.cuda.synchronize()
index.check()
msp = torch.tensor([e0.elapsed_time(e1) / STEPS], device="cuda", dtype=torch.float64)
dist.all_reduce(msp, op=dist.ReduceOp.MAX)
if rank == 0:
    print("peer-memory exchange: sustained %.3f ms/search = %.0f q/s, identical to the NCCL path: %s" % (msp.item(), Q / msp.item() * 1e3, same), flush=True)
if rank == 0:
    print("world %d rows/rank %d: " % (world, s1 - s0) + "  ".join("%s %.3f" % (n, v) for n, v in zip(names + ["sum"], t.tolist())) +
          " ms | sustained %.3f ms/search = %.0f q/s | stats %s" % (ms.item(), Q / ms.item() * 1e3, index.local.stats()), flush=True)
index.disable_peer_exchange()
dist.destroy_process_group()
