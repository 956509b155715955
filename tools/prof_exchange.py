#!/usr/bin/env python3
"""Times the pieces of the multi-GPU commit separately (torchrun, N ranks): LDE into the local send buffer, LDE with
fused peer stores, NCCL all-to-all, leaf hashing + tree.  Development tool."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import torch.distributed as dist
import eth_lc_plonky2_b200 as E
from eth_lc_plonky2_b200 import _lib
from eth_lc_plonky2_b200._lib import check

rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
E.init(local_rank)
stream = torch.cuda.Stream()
E.set_stream(stream.cuda_stream)
cols, log_n, r, h = 135, 20 + (world.bit_length() - 1), 3, 4
n = 1 << log_n
plan = E.ShardPlan(cols, log_n, r, h, world)
my = plan.columns_of(rank)
dev = torch.from_numpy(E.splitmix_columns(len(my), n, first_col=my.start).view(np.int64)).cuda()
ex = E.PeerExchange(plan, rank, torch.device("cuda", local_rank))
ops = E.EngineOps(torch.device("cuda", local_rank))
c_r = len(my)
coeffs = ops.empty(c_r * n).view(c_r, n)
send = ops.empty(c_r * (n << r))
recv = ops.empty(cols * plan.rows_per_rank)


def timed(name, fn, reps=4):
    ts = []
    for i in range(reps):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            fn()
            e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = torch.tensor([min(ts[1:])], device="cuda")
    allt = [torch.zeros(1, device="cuda") for _ in range(world)]
    dist.all_gather(allt, t)
    if rank == 0:
        per = [x.item() for x in allt]
        print("%-46s %8.2f ms (max over ranks, best of %d)  per rank: %s" % (name, max(per), reps - 1, " ".join("%.1f" % v for v in per)), flush=True)


lib = _lib.lib()
timed("LDE -> local send buffer [G][C_r][L/G]", lambda: check(lib.eng_lde_dev(C.c_void_p(dev.data_ptr()), c_r, log_n, r, 1, plan.log_world,
                                                                                C.c_void_p(coeffs.data_ptr()), C.c_void_p(send.data_ptr()))))
for chunk_cols in (0, 1, 2, 4, 8):
    E.set_option("lde_peer_chunk_cols", chunk_cols)
    timed("LDE with fused peer stores, chunk_cols=%d" % chunk_cols, lambda: check(lib.eng_lde_peer_dev(C.c_void_p(dev.data_ptr()), c_r, log_n, r, 1, plan.log_world,
                                                                                                      C.c_void_p(coeffs.data_ptr()), C.c_void_p(send.data_ptr()), ex.shard_out, rank)))
E.set_option("lde_peer_chunk_cols", 0)
with torch.cuda.stream(stream):
    timed("NCCL all_to_all_single", lambda: dist.all_to_all_single(recv, send, output_split_sizes=plan.recv_splits(), input_split_sizes=plan.send_splits(rank)))


def tree():
    t = ops.merkle(ex.recv, cols, plan.rows_per_rank, plan.local_cap_height)
    return t


timed("leaf hashing + digest levels (row shard)", tree)
timed("barrier (all_reduce + sync)", ex.barrier)
ex.close()
dist.destroy_process_group()
