#!/usr/bin/env python3
"""Synthetic circuit-shaped proof on the device: prof_prove.py [degree_bits] [reps]."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import eth_lc_plonky2_b200 as E

db = int(sys.argv[1]) if len(sys.argv) > 1 else 18
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
E.init(0)
t0 = time.time()
s = E.synth_circuit(db, seed=1)
t1 = time.time()
circ = E.Circuit.build(s)
E.synchronize()
t2 = time.time()
print("degree_bits %d: synth %.1fs, build (constants||sigmas commit, 84 cols) %.2fs" % (db, t1 - t0, t2 - t1))
wires = list(s["wires"])
for r in range(reps):
    t = time.time()
    proof, ms = circ.prove(wires, s["pi_hash"])
    wall = time.time() - t
    print("prove wall %.3fs  proof %d u64  stages(ms): %s" % (wall, proof.size, {k: round(v, 1) for k, v in ms.items()}))
