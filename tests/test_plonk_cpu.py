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


@pytest.mark.parametrize("native", [3, 1, 0], ids=["compiled+poseidon_fp64", "bytecode+poseidon_fp64", "all_bytecode"])
@pytest.mark.parametrize("db", [3, 5])
def test_quotient_point_replay(emu, oracle, synth, db, native):
    """The four quotient kernels' bodies (PoseidonGate through the FP64 evaluator or through its bytecode; the other gates
    through the interpreter) against the oracle's formulas."""
    _quotient_replay(emu, oracle, synth[db], db, native)


@pytest.mark.parametrize("log_shards", [1, 2, 3])
def test_row_sharded_quotient_replay(emu, oracle, synth, log_shards):
    """The multi-GPU prover's quotient: every row shard evaluated from its own leaf matrices (pos0 / count / by_position),
    gathered and un-sharded, equals the single-device values (Z(w_n x) stays inside a shard for up to 2^3 shards)."""
    s = synth[5]
    circ = oracle.Circuit(s["blob"])
    rng = np.random.default_rng(77)
    betas, gammas, alphas = rand_field(rng, 2), rand_field(rng, 2), rand_field(rng, 2)
    cs = oracle.Batch.from_values(np.concatenate([s["constants"], s["sigmas"]]), 3, 4)
    wires = oracle.Batch.from_values(s["wires"], 3, 4)
    zs = oracle.Batch.from_values(circ.partial_products(s["wires"], s["sigmas"], betas, gammas), 3, 4)
    lde = lambda b: np.ascontiguousarray(b.leaves.T)
    L = 8 << 5
    args = (s["blob"], lde(cs), lde(wires), lde(zs), s["pi_hash"], betas, gammas, alphas)
    emu.emu_quotient_values_sharded.argtypes = [u64p] * 9 + [C.c_int, C.c_uint32]
    one, many = np.zeros((2, L), np.uint64), np.zeros((2, L), np.uint64)
    assert emu.emu_quotient_values_sharded(*args, one, 1, 0) == 0
    assert emu.emu_quotient_values_sharded(*args, many, 1, log_shards) == 0
    assert (one == many).all() and one.any()
    assert emu.emu_quotient_values_sharded(*args, many, 1, 4) == 2       # 16 shards > 2^quotient_degree_bits: refused


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


@pytest.mark.parametrize("native", [3, 2, 1, 0], ids=["compiled+poseidon_fp64", "compiled+poseidon_bytecode", "bytecode+poseidon_fp64", "all_bytecode"])
def test_all_gate_kinds_quotient_replay(emu, oracle, synth_v2, native):
    """Bytecode (gate_lib.h through the interpreter), the COMPILED evaluators (the same gate_lib.h source instantiated with the
    in-place builder model) and formulas (oracle/gates.h) for all 22 gate kinds, through the quotient kernels' bodies."""
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


# ---- row f4 on CPU: the product's verifier and byte format are host code (no device), checked against the ORACLE's prover ----
@pytest.mark.parametrize("which,db", [("v1", 5), ("v1", 7), ("v2", 5), ("v2", 6)])
def test_engine_verifier_accepts_oracle_proofs_and_rejects_tampering(oracle, synth, synth_v2, which, db):
    import eth_lc_plonky2_b200 as E
    s = (synth if which == "v1" else synth_v2)[db]
    circ = oracle.Circuit(s["blob"])
    cs = oracle.Batch.from_values(np.concatenate([s["constants"], s["sigmas"]]), 3, 4)
    proof = circ.prove(cs, s["wires"], s["sigmas"], s["pi_hash"])
    E.verify(s["blob"], cs.cap, s["pi_hash"], proof)                          # two verifiers, one proof
    assert circ.verify(cs.cap, s["pi_hash"], proof) == 0
    rng = np.random.default_rng(db)
    for k in rng.integers(0, proof.size, 24):                                  # any flipped word is caught (or unparsable)
        bad = proof.copy(); bad[int(k)] ^= np.uint64(1)
        with pytest.raises(E.EngineError):
            E.verify(s["blob"], cs.cap, s["pi_hash"], bad)
    pi2 = s["pi_hash"].copy(); pi2[3] ^= np.uint64(1)
    with pytest.raises(E.EngineError, match="vanishing"):
        E.verify(s["blob"], cs.cap, pi2, proof)
    bad_cap = np.array(cs.cap).copy(); bad_cap[0, 0] ^= np.uint64(1)
    with pytest.raises(E.EngineError, match="Merkle"):
        E.verify(s["blob"], bad_cap, s["pi_hash"], proof)
    # an unsatisfied constraint: the prover still emits a proof, both verifiers reject at the vanishing identity
    row = int(np.where((s["constants"][:1] != 0xFFFFFFFF).any(axis=0))[0][1])
    w = s["wires"].copy(); w[0, row] ^= np.uint64(1)
    badp = circ.prove(cs, w, s["sigmas"], s["pi_hash"])
    if circ.verify(cs.cap, s["pi_hash"], badp) != 0:
        with pytest.raises(E.EngineError):
            E.verify(s["blob"], cs.cap, s["pi_hash"], badp)


@pytest.mark.parametrize("which,db", [("v1", 5), ("v2", 6)])
def test_proof_bytes_round_trip(oracle, synth, synth_v2, which, db):
    """ProofWithPublicInputs::to_bytes / from_bytes: the byte length follows from the circuit's shapes (8 bytes per field
    element, one count byte per Merkle path), the round trip is the identity and the reparsed proof still verifies."""
    import eth_lc_plonky2_b200 as E
    s = (synth if which == "v1" else synth_v2)[db]
    circ = oracle.Circuit(s["blob"])
    cs = oracle.Batch.from_values(np.concatenate([s["constants"], s["sigmas"]]), 3, 4)
    proof = circ.prove(cs, s["wires"], s["sigmas"], s["pi_hash"])
    pis = [1, 2, 3, 0xFFFFFFFF00000000]
    data = E.proof_to_bytes(s["blob"], proof, pis)
    nconst = s["constants"].shape[0]
    arity = oracle.fri_arity_bits(db)
    L_bits = db + 3
    words = 3 * 64 + 2 * (nconst + 80 + 135 + 2 + 2 + 18 + 16) + len(arity) * 64 + 2 * (1 << (db - 4 * len(arity))) + 1 + len(pis)
    paths = 0
    for _ in range(28):
        words += (nconst + 80) + 135 + 20 + 16 + 4 * 4 * (L_bits - 4)
        paths += 4
        lb = L_bits
        for a in arity:
            lb -= a
            words += 2 * 16 + 4 * max(lb - 4, 0)
            paths += 1
    assert len(data) == 8 * words + paths
    back, pis2 = E.proof_from_bytes(s["blob"], data)
    assert (back == proof).all() and list(pis2) == pis
    E.verify(s["blob"], cs.cap, s["pi_hash"], back)
    with pytest.raises(E.EngineError):
        E.proof_from_bytes(s["blob"], data[:-3])
    nc = bytearray(data); nc[0:8] = (0xFFFFFFFFFFFFFFFF).to_bytes(8, "little")   # a non-canonical element is refused
    with pytest.raises(E.EngineError):
        E.proof_from_bytes(s["blob"], bytes(nc))


def test_circuit_descriptions_are_validated():
    """Programs that arrive over the ABI are checked before they reach a kernel: wire / constant / immediate indices,
    registers read before written, unknown opcodes, constraint counts."""
    import eth_lc_plonky2_b200 as E
    s = E.synth_circuit_v2(4, seed=3, kinds_mask=(1 << 3) | (1 << 5))
    blob = s["blob"]
    cap = np.zeros((16, 4), np.uint64)
    ng = int(blob[2 + 11])
    prog0 = 20 + 12 * ng
    def rejected(b, pat):
        with pytest.raises(E.EngineError, match=pat):
            E.verify(b, cap, s["pi_hash"], np.zeros(8, np.uint64))
    b = blob.copy(); b[prog0] = np.uint64(99); rejected(b, "unknown opcode|invalid program")
    b = blob.copy(); b[prog0] = np.uint64(int(b[prog0]) | (0x1FFF << 20)); rejected(b, "invalid program")
    b = blob.copy(); b[1] += np.uint64(1); rejected(b, "inconsistent")
    b = blob.copy(); b[20 + 1] = np.uint64(77); rejected(b, "selector index")
    rejected(blob, "does not parse")                                          # a well-formed circuit, a garbage proof


# ---- second restatement (pure Python, tests/golden/plonk_restatement.py) pins oracle/plonk.h + oracle/fri.h ----
def _golden_plonk_cases():
    import json, os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "plonk_proof.json")) as f:
        return json.load(f)["cases"]


def golden_plonk_circuit(case):
    import eth_lc_plonky2_b200 as E
    s = E.synth_circuit(case["degree_bits"], seed=case["seed"])
    s["blob"] = s["blob"].copy()
    s["blob"][9], s["blob"][10] = case["pow_bits"], case["num_query_rounds"]
    return s, np.array([int(x, 16) for x in case["proof"]], np.uint64)


@pytest.mark.parametrize("k", [0, 1])
def test_oracle_proof_equals_python_restatement(oracle, k):
    """The C++ oracle (FFT butterflies, fast partial rounds, Horner division, fill_subtree) and the pure-Python restatement
    (direct DFT sums, naive rounds, synthetic division in coefficient space, level-by-level trees) give the SAME proof, word
    for word: caps, openings, FRI commit phase, grind, query openings.  Both verifiers and the product verifier accept it."""
    import hashlib
    import eth_lc_plonky2_b200 as E
    case = _golden_plonk_cases()[k]
    s, gold = golden_plonk_circuit(case)
    circ = oracle.Circuit(s["blob"])
    cs = oracle.Batch.from_values(np.concatenate([s["constants"], s["sigmas"]]), 3, 4)
    assert [int(x) for x in np.array(cs.cap).ravel()] == case["cs_cap"]
    betas, gammas = np.array(case["betas"], np.uint64), np.array(case["gammas"], np.uint64)
    zs = circ.partial_products(s["wires"], s["sigmas"], betas, gammas)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).astype("<u8").tobytes()).hexdigest()
    assert sha(zs) == case["sha256_zs_pp"]
    wires = oracle.Batch.from_values(s["wires"], 3, 4)
    zb = oracle.Batch.from_values(zs, 3, 4)
    q = circ.quotient(cs, wires, zb, s["pi_hash"], betas, gammas, np.array(case["alphas"], np.uint64))
    assert sha(q) == case["sha256_quotient_coeffs"]
    proof = circ.prove(cs, s["wires"], s["sigmas"], s["pi_hash"])
    assert proof.shape == gold.shape and (proof == gold).all()
    assert circ.verify(cs.cap, s["pi_hash"], gold) == 0
    E.verify(s["blob"], cs.cap, s["pi_hash"], gold)


def test_python_verifier_accepts_oracle_proofs_and_rejects_tampering(oracle):
    """The pure-Python verifier on a proof of the C++ oracle for ANOTHER witness (not the golden one)."""
    import importlib.util, os
    import eth_lc_plonky2_b200 as E
    spec = importlib.util.spec_from_file_location("plonk_restatement", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "plonk_restatement.py"))
    R = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(R)
    s = E.synth_circuit(4, seed=99)
    s["blob"] = s["blob"].copy(); s["blob"][9], s["blob"][10] = 4, 2
    circ = oracle.Circuit(s["blob"])
    cs = oracle.Batch.from_values(np.concatenate([s["constants"], s["sigmas"]]), 3, 4)
    proof = [int(x) for x in circ.prove(cs, s["wires"], s["sigmas"], s["pi_hash"])]
    blob, cap, pi = [int(x) for x in s["blob"]], [int(x) for x in np.array(cs.cap).ravel()], [int(x) for x in s["pi_hash"]]
    assert R.verify(blob, cap, pi, proof) is None
    bad = list(proof); bad[700] ^= 1
    assert R.verify(blob, cap, pi, bad) is not None
    assert R.verify(blob, cap, [pi[0] ^ 1] + pi[1:], proof) == "vanishing polynomial identity"


# ---- row f2, host half of build(): sigma polynomials from copy constraints ----
def _sigma_value_table(db, num_routed):
    n = 1 << db
    g = pow(1753635133440165772, 1 << (32 - db), P_)
    xs = [pow(g, i, P_) for i in range(n)]
    return {(7 ** c % P_) * xs[r] % P_: (r, c) for c in range(num_routed) for r in range(n)}


P_ = 0xFFFFFFFF00000001


def test_build_sigmas_reproduces_the_synthetic_circuits(synth):
    """The copy constraints read back out of the synthetic circuit's sigma columns give the same columns again."""
    import eth_lc_plonky2_b200 as E
    s = synth[5]
    table = _sigma_value_table(5, 80)
    copies = []
    for c in range(80):
        for r in range(32):
            r2, c2 = table[int(s["sigmas"][c][r])]
            if (r2, c2) > (r, c):
                copies.append((r, c, r2, c2))
    assert len(copies) > 50
    assert (E.build_sigmas(5, 80, copies) == s["sigmas"]).all()
    assert (E.build_sigmas(5, 80, []) == E.build_sigmas(5, 80, [(3, 4, 3, 4)])).all()          # identity permutation
    with pytest.raises(E.EngineError):
        E.build_sigmas(5, 80, [(0, 80, 1, 0)])


def test_build_sigmas_against_a_python_union_find():
    """Classes of any size (chains, stars, redundant constraints): every class is one cycle in (row, column) order, as
    plonky2's WirePartition lists it; checked against a plain Python partition refinement."""
    import eth_lc_plonky2_b200 as E
    db, nr = 4, 7
    n = 1 << db
    rng = np.random.default_rng(5)
    copies = [tuple(int(v) for v in (rng.integers(n), rng.integers(nr), rng.integers(n), rng.integers(nr))) for _ in range(60)]
    copies += [(1, 1, 2, 2), (2, 2, 3, 3), (3, 3, 1, 1), (5, 0, 5, 0)]                      # a redundant triangle, a self-copy
    cls = {(r, c): {(r, c)} for r in range(n) for c in range(nr)}
    for ra, ca, rb, cb in copies:
        a, b = cls[(ra, ca)], cls[(rb, cb)]
        if a is not b:
            a |= b
            for w in b:
                cls[w] = a
    got = E.build_sigmas(db, nr, copies)
    table = _sigma_value_table(db, nr)
    seen = set()
    for w, members in cls.items():
        if id(members) in seen:
            continue
        seen.add(id(members))
        order = sorted(members)
        for i, (r, c) in enumerate(order):
            assert table[int(got[c][r])] == order[(i + 1) % len(order)]
    assert max(len(m) for m in cls.values()) > 3


def test_public_inputs_hash_is_hash_no_pad(oracle):
    import eth_lc_plonky2_b200 as E
    for n in (0, 1, 8, 16, 19):
        pis = [(i * 0x9E3779B97F4A7C15 + 5) % (2**64) for i in range(n)]
        want = oracle.hash_no_pad(np.array(pis, np.uint64)) if n else np.zeros(4, np.uint64)
        assert (E.public_inputs_hash(pis) == want).all()
