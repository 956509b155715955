"""Rows a5 / a6 on CPU: (1) the oracle's prover against its own restatement of plonky2's verifier on the synthetic
circuit (prove -> verify, like the reference's tests); (2) CPU replay of the CUDA kernel bodies quot_point / pp_row /
pp_finish (tests/emu) against the oracle."""
import ctypes as C

import numpy as np
import pytest

from helpers import rand_field
from test_replay import emu, u64p  # noqa: F401  (fixture)


@pytest.fixture(scope="module")
def synth():
    import eth_lc_plonky2_b200 as E
    return {db: E.synth_circuit(db, seed=11 + db) for db in (3, 5, 7)}


@pytest.mark.parametrize("db", [3, 5, 7])
def test_oracle_prove_then_verify(oracle, synth, db):
    s = synth[db]
    circ = oracle.Circuit(s["blob"])
    cs = oracle.Batch.from_values(np.concatenate([s["constants"], s["sigmas"]]), 3, 4)
    proof = circ.prove(cs, s["wires"], s["sigmas"], s["pi_hash"])
    assert circ.verify(cs.cap, s["pi_hash"], proof) == 0
    # an unsatisfied gate constraint, a broken copy constraint and a wrong public input must all be rejected
    kinds = s["constants"][0]
    arith_row = int(np.where(kinds == 3)[0][0])
    bad = s["wires"].copy(); bad[3, arith_row] ^= np.uint64(1)             # output wire of an arithmetic op
    assert circ.verify(cs.cap, s["pi_hash"], circ.prove(cs, bad, s["sigmas"], s["pi_hash"])) == 21
    noop_rows = np.where(kinds == 0)[0]
    if len(noop_rows) >= 2:
        bad = s["wires"].copy(); bad[0, int(noop_rows[0])] ^= np.uint64(1)   # one end of a copy constraint
        assert circ.verify(cs.cap, s["pi_hash"], circ.prove(cs, bad, s["sigmas"], s["pi_hash"])) == 21
    pi2 = s["pi_hash"].copy(); pi2[0] ^= np.uint64(1)
    assert circ.verify(cs.cap, pi2, proof) != 0
    tampered = proof.copy(); tampered[5] ^= np.uint64(1)                     # a cap element
    assert circ.verify(cs.cap, s["pi_hash"], tampered) != 0


@pytest.mark.parametrize("db", [3, 5, 7])
def test_partial_products_replay(emu, oracle, synth, db):
    emu.emu_partial_products.argtypes = [u64p] * 6
    s = synth[db]
    circ = oracle.Circuit(s["blob"])
    rng = np.random.default_rng(db)
    betas, gammas = rand_field(rng, 2), rand_field(rng, 2)
    ref = circ.partial_products(s["wires"], s["sigmas"], betas, gammas)
    out = np.zeros_like(ref)
    emu.emu_partial_products(s["blob"], s["wires"], s["sigmas"], betas, gammas, out)
    assert (out == ref).all()
    n = 1 << db
    # the copy constraints are satisfied, so Z wraps around to 1: Z(x_{n-1}) * row product = 1 (checked via Z_0 column)
    assert (ref[0:2, 0] == 1).all()


@pytest.fixture(scope="module")
def emu_int_gate():
    """The replay harness built with -DPLK_POSEIDON_F64=0: the PoseidonGate evaluator on the integer pipes ("fast" partial
    rounds), kept as the A/B baseline of the default FP64 evaluator."""
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    src = os.path.join(here, "emu", "emu.cpp")
    so = os.path.join(here, "emu", "libemu_intgate.so")
    csrc = os.path.join(here, "..", "eth-lc-plonky2_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DPLK_POSEIDON_F64=0", "-o", so, src])
    return C.CDLL(so)


@pytest.mark.parametrize("variant", ["default", "int_gate"])
@pytest.mark.parametrize("db", [3, 5])
def test_quotient_point_replay(emu, emu_int_gate, oracle, synth, db, variant):
    if variant == "int_gate":
        emu = emu_int_gate
    emu.emu_quotient_values.argtypes = [u64p] * 9
    s = synth[db]
    circ = oracle.Circuit(s["blob"])
    rng = np.random.default_rng(50 + db)
    betas, gammas, alphas = rand_field(rng, 2), rand_field(rng, 2), rand_field(rng, 2)
    cs = oracle.Batch.from_values(np.concatenate([s["constants"], s["sigmas"]]), 3, 4)
    wires = oracle.Batch.from_values(s["wires"], 3, 4)
    zs = oracle.Batch.from_values(circ.partial_products(s["wires"], s["sigmas"], betas, gammas), 3, 4)
    ref = circ.quotient(cs, wires, zs, s["pi_hash"], betas, gammas, alphas)        # [16][n] coefficient chunks
    L = 8 << db
    vals = np.zeros((2, L), np.uint64)
    lde = lambda b: np.ascontiguousarray(b.leaves.T)                               # engine layout: [cols][L] bit-reversed rows
    emu.emu_quotient_values(s["blob"], lde(cs), lde(wires), lde(zs), s["pi_hash"], betas, gammas, alphas, vals)
    # coset_ifft(7) of the replayed values must give the oracle's coefficients
    for c in range(2):
        co = oracle.ifft(vals[c])
        inv7 = oracle.lib().orc_gl_inv(7)
        pw = np.array([pow(inv7, k, oracle.P) for k in range(L)], dtype=object)
        co = np.array([(int(a) * int(b)) % oracle.P for a, b in zip(co, pw)], dtype=np.uint64)
        assert (co.reshape(8, -1) == ref[8 * c:8 * c + 8]).all()
