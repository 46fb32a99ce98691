#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU): row-sharded search + one all-gather + merge and
sharded alpha-QE against the single-process CPU oracle."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch
import torch.distributed as dist

import synthdata as synth
from dirb200.dist import ShardedIndex, shard_rows
from oracle import dir_oracle as O

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
db, q, pos = synth.make_descriptor_db(40000, 70, dim=2048, n_pos=10)
s0, s1 = shard_rows(db.shape[0], world, rank)
index = ShardedIndex(torch.from_numpy(db[s0:s1]).cuda(), row_offset=s0)
index.local.set_option("sample_rows", 2048)
qd = torch.from_numpy(q).cuda()
s, i = index.search(qd, 100)
ok = True
if rank == 0:
    rs, ri = O.topk(q, db, 100)
    ok = bool(np.array_equal(i.cpu().numpy(), ri) and np.abs(s.cpu().numpy() - rs).max() < 1e-12)
out = index.expand_queries(qd, 2, 0.5)
if rank == 0:
    ref = O.expand_descriptors(q, db=db, k=2, alpha=0.5)
    ok = ok and float(np.linalg.norm(out.cpu().numpy() - ref) / np.linalg.norm(ref)) < 1e-5
# exact-mAP rank counting over the sharded database: scores from the owning shard, counts summed over shards
gnd = synth.oxford_gt(pos, n_junk=3, n_db=db.shape[0], seed=5)
offs, rows, flags = [0], [], []
for g_ in gnd:
    r_ = sorted(set(g_["ok"]) | set(g_["junk"]))
    rows += r_
    flags += [1 if x in set(g_["ok"]) else 0 for x in r_]
    offs.append(len(rows))
sc, above = index.rank_counts(qd, np.array(offs, np.int32), np.array(rows, np.int64), np.array(flags, np.uint8))
if rank == 0:
    rs_, ra_ = O.rank_counts(q, db, offs, np.array(rows))
    m_ = np.array(flags) == 1
    ok = ok and bool(np.abs(sc.cpu().numpy() - rs_).max() < 1e-12 and np.array_equal(above.cpu().numpy()[m_], ra_[m_]))
# Persisted path (opt-in: DIST_CHECK_STORE=<directory on a filesystem all ranks see>): every rank writes its rows as a
# shard of a descriptor store, the store is re-read under the current world size (row ranges come from the manifest)
# and searched again.
if os.environ.get("DIST_CHECK_STORE"):
    from dirb200 import store as S
    st = S.write_distributed(os.environ["DIST_CHECK_STORE"], db[s0:s1])
    index2 = ShardedIndex.from_store(st, "cuda:%d" % local)
    s2, i2 = index2.search(qd, 100)
    ok = ok and bool(torch.equal(i2, i) and torch.equal(s2, s))
# The same searches through the peer-memory exchange (no NCCL on the search path): thresholds and lists are stored by the
# search kernels into every rank's window; every rank must end up with exactly the NCCL-path result, search after search.
index.enable_peer_exchange(max_q=128, max_k=128)
for _ in range(4):
    sp, ip = index.search(qd, 100)
    ok = ok and bool(torch.equal(ip, i) and torch.equal(sp, s))
outp = index.expand_queries(qd, 2, 0.5)
ok = ok and bool(torch.equal(outp, out))
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DIST-OK" if int(flag.item()) == 1 else "DIST-FAIL", "world", world)
index.disable_peer_exchange()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
