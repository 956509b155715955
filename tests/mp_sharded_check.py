#!/usr/bin/env python3
"""Worker of tests/test_gpu_multiprocess.py (launched with torch.distributed.run, one process per GPU): the multi-GPU commit
and the sharded prover over the REAL transports -- NCCL collectives and CUDA-IPC peer stores -- against the single-GPU path
of rank 0.  Exits non-zero on any mismatch."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import torch.distributed as dist
import eth_lc_plonky2_b200 as E

rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
E.init(local_rank)
dev = torch.device("cuda", local_rank)
ok = True


def check(name, cond):
    global ok
    if not cond:
        ok = False
        print("rank %d: MISMATCH %s" % (rank, name), flush=True)


# ---- commit: peer stores and NCCL all-to-all, device and host columns, NO set_stream ----
C_, log_n, r, h = 37, 12, 3, 4
full = E.splitmix_columns(C_, 1 << log_n)
plan = E.ShardPlan(C_, log_n, r, h, world)
cols = plan.columns_of(rank)
single = E.PolynomialBatch.from_values(list(full), r, False, h) if rank == 0 else None
ex = E.PeerExchange(plan, rank, dev, slots=2)
local = torch.from_numpy(full[cols.start:cols.stop].view(np.int64)).to(dev)
for label, kwargs, src in (("peer/device", dict(exchange=ex, slot=0), local), ("peer/host", dict(exchange=ex, slot=1), [full[c] for c in cols]),
                           ("nccl/device", dict(), local)):
    b = E.ShardedPolynomialBatch.from_values(src, plan, rank, **kwargs)
    digs = [None] * world
    dist.all_gather_object(digs, np.ascontiguousarray(b.local_digests))
    if rank == 0:
        check(label + " cap", (np.array(b.cap) == single.merkle_tree.cap).all())
        check(label + " digests", (np.concatenate(digs) == single.merkle_tree.digests).all())
    k = rank * plan.rows_per_rank + 3
    rows = [None] * world
    dist.all_gather_object(rows, (k, b.get(k), b.prove(k)))
    if rank == 0:
        for kk, leaf, path in rows:
            check(label + " leaf %d" % kk, (leaf == single.merkle_tree.get(kk)).all() and (path == single.merkle_tree.prove(kk)).all())
first = E.ShardedPolynomialBatch.from_values(local, plan, rank, exchange=ex, slot=0)
second = E.ShardedPolynomialBatch.from_values(local, plan, rank, exchange=ex, slot=0)      # reuses slot 0: `first` is dead
try:
    first.rows
    check("stale slot detection", False)
except E.EngineError:
    pass
del first, second
ex.close()

# ---- one proof over all ranks, both gate sets ----
for which, db in (("v1", 10), ("v2", 9)):
    s = E.synth_circuit(db, seed=7) if which == "v1" else E.synth_circuit_v2(db, seed=7)
    pr = E.ShardedProver(s["blob"], s["constants"], s["sigmas"], rank, world, device=dev)
    proof, _ = pr.prove(s["wires"], s["pi_hash"])
    pr.verify(s["pi_hash"], proof)
    shas = [None] * world
    dist.all_gather_object(shas, hashlib.sha256(proof.tobytes()).hexdigest())
    check("proof identical on all ranks (%s)" % which, len(set(shas)) == 1)
    pr.close()
    if rank == 0:
        circ = E.Circuit.build(s)
        ref, _ = circ.prove(s["wires"], s["pi_hash"])
        check("sharded proof == single-GPU proof (%s)" % which, ref.shape == proof.shape and (ref == proof).all())
flag = torch.tensor([0 if ok else 1], device=dev)
dist.all_reduce(flag)
dist.destroy_process_group()
if rank == 0:
    print("mp_sharded_check: %s (world %d)" % ("OK" if flag.item() == 0 else "FAILED", world), flush=True)
sys.exit(0 if flag.item() == 0 else 1)
