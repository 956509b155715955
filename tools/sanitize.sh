#!/bin/bash
# compute-sanitizer runs (VERDICT r1 task 10): memcheck + racecheck over a small commit and a small proof,
# logs under gpurun_out/ (copied to profiles/ by hand).  Usage on the GPU box:  bash tools/sanitize.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
import eth_lc_plonky2_b200 as E
E.init(0)
vals = E.splitmix_columns(135, 1 << 12)
b = E.PolynomialBatch.from_values(list(vals), 3, False, 4)
print("commit 135 x 2^12 cap0 %016x" % int(b.merkle_tree.cap[0][0]))
b.merkle_tree.prove(77); b.merkle_tree.get(5)
bb = E.PolynomialBatch.from_values(list(vals[:9]), 2, True, 1, blinding_seed=5)
s = E.synth_circuit_v2(8, seed=3)
c = E.Circuit.build(s)
p, _ = c.prove(s["wires"], s["pi_hash"])
c.verify(s["pi_hash"], p)
print("proof 2^8 rows, 18 gate kinds: %d words, verified" % p.size)
pr = E.ShardedProver(s["blob"], s["constants"], s["sigmas"], 0, 1, use_peer=False)
q, _ = pr.prove(s["wires"], s["pi_hash"])
assert (q == p).all()
print("sharded prover (world 1) identical")
PY
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_case.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "$tool exit code $?" >> gpurun_out/r02_sanitizer_$tool.log
  tail -4 gpurun_out/r02_sanitizer_$tool.log
done
