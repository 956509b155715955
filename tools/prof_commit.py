#!/usr/bin/env python3
"""One PolynomialBatch::from_values on the device (for ncu / quick timing): prof_commit.py [log_n] [cols] [reps] [lde_group_mb]."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import eth_lc_plonky2_b200 as E

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 17
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 135
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
group_mb = int(sys.argv[4]) if len(sys.argv) > 4 else None
E.init(0)
if group_mb is not None:
    E.set_option("lde_group_mb", group_mb)
vals = torch.from_numpy(E.splitmix_columns(cols, 1 << log_n).view(np.int64)).cuda()
for _ in range(reps):
    b = E.PolynomialBatch.from_values(vals, 3, False, 4)
    ms = b.stage_ms()
    cap0 = "%016x" % int(b.merkle_tree.cap[0][0])
    b.close()
print("lde_group_mb", group_mb, "log_n", log_n, "cols", cols, "cap0", cap0, {k: round(v, 3) for k, v in ms.items()})
