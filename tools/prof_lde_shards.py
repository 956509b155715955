#!/usr/bin/env python3
"""One rank's transform of the multi-GPU commit on ONE GPU: iNTT + LDE of `cols` columns x 2^log_n rows written as 2^log_shards
row shards [G][C][L/G] (eng_lde_dev).  prof_lde_shards.py [log_n] [cols] [log_shards] [reps]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import eth_lc_plonky2_b200 as E
from eth_lc_plonky2_b200._lib import check, lib

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 23
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 17
log_shards = int(sys.argv[3]) if len(sys.argv) > 3 else 3
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
E.init(0)
stream = torch.cuda.Stream()
E.set_stream(stream.cuda_stream)
n = 1 << log_n
dev = torch.from_numpy(E.splitmix_columns(cols, n).view(np.int64)).cuda()
coeffs = torch.empty(cols * n, dtype=torch.int64, device="cuda")
send = torch.empty(cols * (n << 3), dtype=torch.int64, device="cuda")
ts = []
for _ in range(reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        check(lib().eng_lde_dev(C.c_void_p(dev.data_ptr()), cols, log_n, 3, 1, log_shards, C.c_void_p(coeffs.data_ptr()), C.c_void_p(send.data_ptr())))
        e1.record(stream)
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("so %s  log_n %d cols %d shards 2^%d: iNTT + LDE %s ms (best %.2f)  checksum %016x" % (
    os.path.basename(E.so_path()), log_n, cols, log_shards, " ".join("%.2f" % t for t in ts), min(ts[1:]), int(send[::4097].sum().item()) & (2**64 - 1)))
