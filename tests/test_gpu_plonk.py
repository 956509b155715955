"""Rows a5 / a6 and the whole prove() on the GPU against the oracle, bit for bit, on the synthetic circuit; every
engine proof is verified by the oracle's restatement of plonky2's verifier (prove -> verify, as the reference's tests)."""
import numpy as np
import pytest

from helpers import rand_field

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("db", [3, 5, 8, 11])
def test_partial_products_match_oracle(engine, oracle, db):
    s = engine.synth_circuit(db, seed=3 + db)
    circ = engine.Circuit.build(s)
    oc = oracle.Circuit(s["blob"])
    rng = np.random.default_rng(db)
    betas, gammas = rand_field(rng, 2), rand_field(rng, 2)
    got = circ.partial_products(s["wires"], betas, gammas)
    assert (got == oc.partial_products(s["wires"], s["sigmas"], betas, gammas)).all()


@pytest.mark.parametrize("db", [3, 6, 9])
def test_quotient_matches_oracle(engine, oracle, db):
    E = engine
    s = E.synth_circuit(db, seed=30 + db)
    circ = E.Circuit.build(s)
    oc = oracle.Circuit(s["blob"])
    rng = np.random.default_rng(60 + db)
    betas, gammas, alphas = rand_field(rng, 2), rand_field(rng, 2), rand_field(rng, 2)
    zvals = oc.partial_products(s["wires"], s["sigmas"], betas, gammas)
    wires = E.PolynomialBatch.from_values(list(s["wires"]), 3, False, 4)
    zs = E.PolynomialBatch.from_values(list(zvals), 3, False, 4)
    q = circ.quotient(wires, zs, s["pi_hash"], betas, gammas, alphas)
    ocs = oracle.Batch.from_values(np.concatenate([s["constants"], s["sigmas"]]), 3, 4)
    ref = oc.quotient(ocs, oracle.Batch.from_values(s["wires"], 3, 4), oracle.Batch.from_values(zvals, 3, 4), s["pi_hash"], betas, gammas, alphas)
    assert q.num_polys == 16 and (q.polynomials == ref).all()
    assert (q.merkle_tree.cap == oracle.Batch.from_coeffs(ref, 3, 4).cap).all()


@pytest.mark.parametrize("db", [3, 5, 8, 10, 12])
def test_prove_matches_oracle_and_verifies(engine, oracle, db):
    E = engine
    s = E.synth_circuit(db, seed=100 + db)
    circ = E.Circuit.build(s)
    proof, ms = circ.prove(s["wires"], s["pi_hash"])
    oc = oracle.Circuit(s["blob"])
    ocs = oracle.Batch.from_values(np.concatenate([s["constants"], s["sigmas"]]), 3, 4)
    assert (circ.constants_sigmas.merkle_tree.cap == ocs.cap).all()
    assert oc.verify(ocs.cap, s["pi_hash"], proof) == 0                 # plonky2's verifier accepts the GPU proof
    ref = oc.prove(ocs, s["wires"], s["sigmas"], s["pi_hash"])
    assert proof.shape == ref.shape and (proof == ref).all()            # caps, openings, FRI: identical to the CPU prover
    assert ms["total"] > 0


def test_invalid_witness_gives_a_proof_that_does_not_verify(engine, oracle):
    """plonky2 does not notice an unsatisfied constraint while proving; verify() rejects (the #[should_panic] behaviour
    of /root/reference/eth-lc-plonky2/src/unit_tests.rs:377 relies on that)."""
    E = engine
    s = E.synth_circuit(6, seed=9)
    circ = E.Circuit.build(s)
    bad = s["wires"].copy()
    row = int(np.where(s["constants"][0] == 3)[0][0])
    bad[7, row] ^= np.uint64(1)
    proof, _ = circ.prove(bad, s["pi_hash"])
    oc = oracle.Circuit(s["blob"])
    assert oc.verify(circ.constants_sigmas.merkle_tree.cap, s["pi_hash"], proof) == 21


# ---- gates as bytecode (row a6 for the real gate set), product verifier (f4), circuit cache (f2) ----
@pytest.mark.parametrize("db", [5, 8, 11])
def test_all_gate_kinds_prove_matches_oracle_and_verifies(engine, oracle, db):
    """22 gate kinds (core, recursion, plonky2_crypto u32) in one circuit: the engine interprets the gates' bytecode, the
    oracle evaluates its own formulas (oracle/gates.h); proofs identical word for word, both verifiers accept."""
    E = engine
    s = E.synth_circuit_v2(db, seed=200 + db)
    circ = E.Circuit.build(s)
    assert circ.info.num_gates == len(E.GATE_KINDS) == 22
    proof, ms = circ.prove(s["wires"], s["pi_hash"])
    oc = oracle.Circuit(s["blob"])
    ocs = oracle.Batch.from_values(np.concatenate([s["constants"], s["sigmas"]]), 3, 4)
    assert oc.verify(ocs.cap, s["pi_hash"], proof) == 0
    circ.verify(s["pi_hash"], proof)
    ref = oc.prove(ocs, s["wires"], s["sigmas"], s["pi_hash"])
    assert proof.shape == ref.shape and (proof == ref).all()


@pytest.mark.parametrize("which,db", [("v1", 7), ("v2", 7)])
def test_native_poseidon_gate_equals_bytecode(engine, which, db):
    """quot_poseidon_kernel (FP64 permutation) and quot_native_kernel (the compiled library evaluators) against the
    interpreter running the same gates' bytecode: all four combinations give the same quotient polynomials."""
    E = engine
    s = E.synth_circuit(db, seed=5) if which == "v1" else E.synth_circuit_v2(db, seed=5)
    circ = E.Circuit.build(s)
    rng = np.random.default_rng(db)
    betas, gammas, alphas = rand_field(rng, 2), rand_field(rng, 2), rand_field(rng, 2)
    wires = E.PolynomialBatch.from_values(list(s["wires"]), 3, False, 4)
    zs = E.PolynomialBatch.from_values(list(circ.partial_products(s["wires"], betas, gammas)), 3, False, 4)
    try:
        outs = []
        for native_poseidon, native_gates in ((1, 1), (1, 0), (0, 1), (0, 0)):    # compiled / FP64 evaluators vs the interpreter
            E.set_option("quot_native_poseidon", native_poseidon)
            E.set_option("quot_native_gates", native_gates)
            outs.append(circ.quotient(wires, zs, s["pi_hash"], betas, gammas, alphas).polynomials)
    finally:
        E.set_option("quot_native_poseidon", 1)
        E.set_option("quot_native_gates", 1)
    a, b = outs[0], outs[3]
    assert all((o == a).all() for o in outs)
    assert (a == b).all()
    assert (a[:, -1] != 0).any()      # degree really reaches 8n - 1 chunks (the quotient is not trivially zero)


def test_engine_verifier_and_bytes_on_gpu_proofs(engine, oracle):
    E = engine
    s = E.synth_circuit_v2(9, seed=77)
    circ = E.Circuit.build(s)
    proof, _ = circ.prove(s["wires"], s["pi_hash"])
    circ.verify(s["pi_hash"], proof)
    data = E.proof_to_bytes(s["blob"], proof, [5, 6])
    back, pis = E.proof_from_bytes(s["blob"], data)
    assert (back == proof).all() and list(pis) == [5, 6]
    bad = proof.copy(); bad[200] ^= np.uint64(1)
    with pytest.raises(E.EngineError):
        circ.verify(s["pi_hash"], bad)
    # an unsatisfied U32 gate constraint: proof is produced, both verifiers reject
    b = [int(x) for x in s["blob"]]
    kinds = [b[20 + 12 * i] for i in range(b[13])]
    gi = kinds.index(E.GATE_KINDS.index("U32Arithmetic"))
    row = int(np.where((s["constants"][:b[6]] == gi).any(axis=0))[0][0])
    w = s["wires"].copy(); w[3, row] ^= np.uint64(1)          # output_low of op 0
    badp, _ = circ.prove(w, s["pi_hash"])
    with pytest.raises(E.EngineError, match="vanishing"):
        circ.verify(s["pi_hash"], badp)
    oc = oracle.Circuit(s["blob"])
    assert oc.verify(circ.constants_sigmas.merkle_tree.cap, s["pi_hash"], badp) == 21


def test_circuit_cache_round_trip(engine, tmp_path):
    """eng_circuit_save / eng_circuit_load (f2): the reloaded prover data gives the same cap and the same proof."""
    E = engine
    s = E.synth_circuit_v2(8, seed=31)
    circ = E.Circuit.build(s)
    path = tmp_path / "circuit.plk"
    circ.save(path)
    again = E.Circuit.load(path)
    assert (again.blob == circ.blob).all()
    assert (again.constants_sigmas.merkle_tree.cap == circ.constants_sigmas.merkle_tree.cap).all()
    p1, _ = circ.prove(s["wires"], s["pi_hash"])
    p2, _ = again.prove(s["wires"], s["pi_hash"])
    assert (p1 == p2).all()
    again.verify(s["pi_hash"], p2)
    raw = bytearray(path.read_bytes())                          # a corrupted coefficient: the rebuilt commitment must not reproduce the stored cap
    blob_words = int(np.frombuffer(bytes(raw[16:24]), np.uint64)[0])
    off = 24 + 8 * blob_words + 16 + 8 * 64 + 8 * 5             # a coefficient of the first constants polynomial
    raw[off] ^= 1
    (tmp_path / "bad.plk").write_bytes(bytes(raw))
    with pytest.raises(E.EngineError, match="cap"):
        E.Circuit.load(tmp_path / "bad.plk")


def test_prove_2_16_matches_oracle(engine, oracle):
    """proof parity at 2^16 rows (the oracle needs about a minute)."""
    E = engine
    s = E.synth_circuit(16, seed=116)
    circ = E.Circuit.build(s)
    proof, _ = circ.prove(s["wires"], s["pi_hash"])
    circ.verify(s["pi_hash"], proof)
    oc = oracle.Circuit(s["blob"])
    ocs = oracle.Batch.from_values(np.concatenate([s["constants"], s["sigmas"]]), 3, 4)
    ref = oc.prove(ocs, s["wires"], s["sigmas"], s["pi_hash"])
    assert proof.shape == ref.shape and (proof == ref).all()


def test_prove_2_18_verifies(engine, oracle):
    """2^18 rows, all gate kinds: verified by the product verifier and by the oracle's (verification is cheap at any size)."""
    E = engine
    s = E.synth_circuit_v2(18, seed=118)
    circ = E.Circuit.build(s)
    proof, _ = circ.prove(s["wires"], s["pi_hash"])
    circ.verify(s["pi_hash"], proof)
    oc = oracle.Circuit(s["blob"])
    assert oc.verify(circ.constants_sigmas.merkle_tree.cap, s["pi_hash"], proof) == 0


@pytest.mark.parametrize("k", [0, 1])
def test_engine_proof_equals_python_restatement(engine, k):
    """eng_prove against the pure-Python second restatement (tests/golden/plonk_proof.json), word for word."""
    from test_plonk_cpu import _golden_plonk_cases, golden_plonk_circuit
    E = engine
    case = _golden_plonk_cases()[k]
    s, gold = golden_plonk_circuit(case)
    circ = E.Circuit.build(s)
    assert [int(x) for x in circ.constants_sigmas.merkle_tree.cap.ravel()] == case["cs_cap"]
    proof, _ = circ.prove(s["wires"], s["pi_hash"])
    assert proof.shape == gold.shape and (proof == gold).all()
