"""a7 + a8 on the GPU against the oracle: opening evaluations and the whole FRI proof, bit for bit.

The oracle proves in plonky2's coefficient space (Horner division, FFT per round); the engine builds the same
codewords point-wise in the evaluation domain, so equality of the proofs is a cross-check of two different algorithms.
Every engine proof is also VERIFIED by the oracle's restatement of plonky2's verifier, as the reference's tests do."""
import numpy as np
import pytest

from helpers import P, rand_field

pytestmark = pytest.mark.gpu


def _setup(engine, oracle, log_n, r, h, widths, seed):
    rng = np.random.default_rng(seed)
    vals = [rand_field(rng, (w, 1 << log_n)) for w in widths]
    eb = [engine.PolynomialBatch.from_values(list(v), r, False, h) for v in vals]
    ob = [oracle.Batch.from_values(v, r, h) for v in vals]
    zeta = tuple(int(x) for x in rand_field(rng, 2))
    g = oracle.lib().orc_root_of_unity(log_n)
    gz = (g * zeta[0] % P, g * zeta[1] % P)
    all_polys = [(o, p) for o, w in enumerate(widths) for p in range(w)]
    nxt = [(2, p) for p in range(min(2, widths[2]))]
    return eb, ob, zeta, gz, [(zeta, all_polys), (gz, nxt)]


@pytest.mark.parametrize("log_n", [1, 5, 8, 9, 12, 14])
def test_eval_ext_matches_oracle(engine, oracle, log_n):
    eb, ob, zeta, gz, _ = _setup(engine, oracle, log_n, 1, 0, [7, 3, 2, 1], log_n)
    for e, o in zip(eb, ob):
        for z in (zeta, gz, (5, 0), (0, 0)):
            assert (e.eval(z) == oracle.batch_eval(o, z)).all()


@pytest.mark.parametrize("log_n,r,h,pow_bits,queries", [(7, 3, 2, 8, 5), (9, 3, 4, 10, 7), (10, 3, 4, 16, 28), (13, 3, 4, 16, 28),
                                                         (14, 1, 0, 12, 9), (9, 2, 1, 9, 4)])
def test_fri_proof_matches_oracle_and_verifies(engine, oracle, log_n, r, h, pow_bits, queries):
    E, O = engine, oracle
    widths = [5, 9, 4, 3]
    eb, ob, zeta, gz, batches = _setup(E, O, log_n, r, h, widths, 100 + log_n)
    arity = O.fri_arity_bits(log_n, r, h)
    assert arity == E.reduction_arity_bits(log_n, r, h)
    ops = np.concatenate([np.concatenate([b.eval(zeta) for b in eb]), eb[2].eval(gz)[:len(batches[1][1])]])
    assert (ops == np.concatenate([np.concatenate([O.batch_eval(b, zeta) for b in ob]), O.batch_eval(ob[2], gz)[:len(batches[1][1])]])).all()
    ech, och = E.Challenger(), O.Challenger()
    for ch in (ech, och):
        (ch.observe_elements if ch is ech else ch.observe)(np.arange(8, dtype=np.uint64))
        (ch.observe_elements if ch is ech else ch.observe)(ops.ravel())
    params = E.FriParams(log_n, r, h, pow_bits, queries)
    proof = E.PolynomialBatch.prove_openings(E.FriInstanceInfo(batches), eb, ech, params)
    oparams = O.fri_params_array(log_n, r, h, pow_bits, queries, arity)
    inst = O.fri_instance_blob(batches)
    ref = O.fri_prove(inst, ob, och, oparams)
    assert proof.blob.shape == ref.shape and (proof.blob == ref).all()       # caps, final poly, PoW, every opening
    assert (ech.state() == och.state()).all()                                 # transcript advanced identically
    vch = O.Challenger(); vch.observe(np.arange(8, dtype=np.uint64)); vch.observe(ops.ravel())
    caps = np.concatenate([b.merkle_tree.cap for b in eb])
    assert O.fri_verify(inst, ops, caps, widths, proof.blob, vch, oparams) == 0   # plonky2's verifier accepts it
    assert len(proof.commit_phase_merkle_caps) == len(arity) and len(proof.query_round_proofs) == queries
    assert proof.final_poly.shape[0] == (1 << (log_n - 4 * len(arity)))
    # a wrong claimed opening must be rejected
    bad = ops.copy(); bad[2, 1] ^= np.uint64(1)
    vch = O.Challenger(); vch.observe(np.arange(8, dtype=np.uint64)); vch.observe(ops.ravel())
    assert O.fri_verify(inst, bad, caps, widths, proof.blob, vch, oparams) != 0


def test_fri_rejects_unsupported(engine):
    E = engine
    b = E.PolynomialBatch.from_values([np.arange(64, dtype=np.uint64)], 3, False, 1)
    inst = E.FriInstanceInfo([((3, 4), [(0, 0)])])
    with pytest.raises(E.EngineError):
        E.PolynomialBatch.prove_openings(inst, [b], E.Challenger(), E.FriParams(6, 3, 1, 8, 2, arity_bits=[3]))
    with pytest.raises(E.EngineError):
        E.PolynomialBatch.prove_openings(E.FriInstanceInfo([((3, 4), [(0, 5)])]), [b], E.Challenger(), E.FriParams(6, 3, 1, 8, 2))
