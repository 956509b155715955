#!/usr/bin/env python3
"""BASELINE.json configs[4]: commit sweep over rows 2^18..2^24 x 135 columns and rate_bits 1..4 on ONE GPU (what fits in
HBM; the rest needs row sharding over several GPUs).  Prints a markdown table: ms per commit (device-resident input, best
of 2 after one warm-up), GB/s of algorithmic bytes B_ntt, stage split.  Development / documentation tool."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import eth_lc_plonky2_b200 as E

COLS, CAP = 135, 4
E.init(0)
free_b, total_b = torch.cuda.mem_get_info()
print("| log2 rows | rate_bits | LDE GB | ms / commit | GB/s (B_ntt) | iNTT | LDE | leaves | levels | cap[0][0] |")
print("|---|---|---|---|---|---|---|---|---|---|")
for log_n in range(18, 25):
    n = 1 << log_n
    vals = torch.from_numpy(E.splitmix_columns(COLS, n).view(np.int64)).cuda()
    for r in (1, 2, 3, 4):
        L = n << r
        need = 8 * COLS * (n + n + L) + 64 * L + (2 << 30)       # values + coeffs + LDE + digests + slack
        if need > free_b * 0.95:
            print("| %d | %d | %.1f | does not fit one GPU (%.0f GB needed) | | | | | | |" % (log_n, r, 8 * COLS * L / 1e9, need / 1e9))
            continue
        best, st, cap0 = None, None, None
        for rep in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            b = E.PolynomialBatch.from_values(vals, r, False, CAP)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if rep and (best is None or ms < best):
                best, st = ms, b.stage_ms()
            cap0 = "%016x" % int(b.merkle_tree.cap[0][0])
            b.close()
        bn = 8 * COLS * n * (2 + (1 << r))
        print("| %d | %d | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %s |" % (
            log_n, r, 8 * COLS * L / 1e9, best, bn / best / 1e6, st["IFFT"], st["FFT + blinding"], st["build Merkle tree (leaves)"],
            st["build Merkle tree (digest levels)"], cap0), flush=True)
    del vals
    torch.cuda.empty_cache()
