#!/usr/bin/env python3
"""Cold start of a prover process (VERDICT r1 task 6): wall clock of eng_init, Circuit.build and the FIRST eng_prove of a
fresh process at 2^log_n rows, next to the second proof.  ENG_TRACE=1 makes the engine print where the first call goes
(pool growth per allocation).   python tools/prof_coldstart.py [log_n] [v1|v2]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
which = sys.argv[2] if len(sys.argv) > 2 else "v1"
t = time.perf_counter()
import eth_lc_plonky2_b200 as E
t_import = time.perf_counter() - t
s = E.synth_circuit(log_n, seed=1) if which == "v1" else E.synth_circuit_v2(log_n, seed=1)
t = time.perf_counter(); E.init(0); t_init = time.perf_counter() - t
t = time.perf_counter(); circ = E.Circuit.build(s); E.synchronize(); t_build = time.perf_counter() - t
wires = list(s["wires"])
walls, stages = [], []
for i in range(3):
    t = time.perf_counter()
    proof, st = circ.prove(wires, s["pi_hash"])
    walls.append(time.perf_counter() - t)
    stages.append({k: round(v, 1) for k, v in st.items()})
circ.verify(s["pi_hash"], proof)
print({"log_n": log_n, "circuit": which, "import_s": round(t_import, 3), "eng_init_s": round(t_init, 3), "build_s": round(t_build, 3),
       "prove_s": [round(w, 3) for w in walls], "stage_ms_first": stages[0], "stage_ms_second": stages[1], "verified": True})
