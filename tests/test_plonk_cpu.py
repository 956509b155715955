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


def _quotient_replay(emu, oracle, s, db, native):
    emu.emu_quotient_values.argtypes = [u64p] * 9 + [C.c_int]
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
    assert emu.emu_quotient_values(s["blob"], lde(cs), lde(wires), lde(zs), s["pi_hash"], betas, gammas, alphas, vals, native) == 0
    # coset_ifft(7) of the replayed values must give the oracle's coefficients
    inv7 = oracle.lib().orc_gl_inv(7)
    pw = np.array([pow(inv7, k, oracle.P) for k in range(L)], dtype=object)
    for c in range(2):
        co = oracle.ifft(vals[c])
        co = np.array([(int(a) * int(b)) % oracle.P for a, b in zip(co, pw)], dtype=np.uint64)
        assert (co.reshape(8, -1) == ref[8 * c:8 * c + 8]).all()


@pytest.mark.parametrize("native", [1, 0], ids=["poseidon_fp64", "poseidon_bytecode"])
@pytest.mark.parametrize("db", [3, 5])
def test_quotient_point_replay(emu, oracle, synth, db, native):
    """The four quotient kernels' bodies (PoseidonGate through the FP64 evaluator or through its bytecode; the other gates
    through the interpreter) against the oracle's formulas."""
    _quotient_replay(emu, oracle, synth[db], db, native)


@pytest.fixture(scope="module")
def synth_v2():
    import eth_lc_plonky2_b200 as E
    return {db: E.synth_circuit_v2(db, seed=5 + db) for db in (5, 6)}


@pytest.mark.parametrize("db", [5, 6])
def test_all_gate_kinds_oracle_prove_then_verify(oracle, synth_v2, db):
    """Every gate of gate_lib.h in one circuit: the oracle evaluates them from its own formulas (oracle/gates.h)."""
    s = synth_v2[db]
    circ = oracle.Circuit(s["blob"])
    cs = oracle.Batch.from_values(np.concatenate([s["constants"], s["sigmas"]]), 3, 4)
    proof = circ.prove(cs, s["wires"], s["sigmas"], s["pi_hash"])
    assert circ.verify(cs.cap, s["pi_hash"], proof) == 0


@pytest.mark.parametrize("native", [1, 0], ids=["poseidon_fp64", "poseidon_bytecode"])
def test_all_gate_kinds_quotient_replay(emu, oracle, synth_v2, native):
    """Bytecode (gate_lib.h) vs formulas (oracle/gates.h) for all 18 gate kinds, through the quotient kernels' bodies."""
    _quotient_replay(emu, oracle, synth_v2[5], 5, native)


def test_every_gate_kind_is_violated_by_a_wrong_wire(oracle, synth_v2):
    """Per gate kind: flipping one constrained wire of a row of that gate makes the oracle's verifier reject (the
    constraints are not vacuous)."""
    import eth_lc_plonky2_b200 as E
    s = synth_v2[5]
    b = [int(x) for x in s["blob"]]
    ng = b[2 + 11]
    kinds = [b[20 + 12 * i] for i in range(ng)]
    circ = oracle.Circuit(s["blob"])
    cs = oracle.Batch.from_values(np.concatenate([s["constants"], s["sigmas"]]), 3, 4)
    nsel = b[2 + 4]
    for gi, kind in enumerate(kinds):
        if E.GATE_KINDS[kind] == "Noop":
            continue
        rows = np.where((s["constants"][:nsel] == gi).any(axis=0))[0]
        assert len(rows), "gate %s has no row" % E.GATE_KINDS[kind]
        bad = s["wires"].copy()
        bad[0, int(rows[0])] ^= np.uint64(1)        # wire 0 is constrained in every gate of the library
        assert circ.verify(cs.cap, s["pi_hash"], circ.prove(cs, bad, s["sigmas"], s["pi_hash"])) != 0, E.GATE_KINDS[kind]
