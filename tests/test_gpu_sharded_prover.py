"""Multi-GPU prover (row e of SURVEY.md 8): ShardedProver against the single-GPU eng_prove, bit for bit.

On the single-GPU test box the ranks run as threads of this process over a thread-based stand-in for torch.distributed
(tests/helpers.py::ThreadDist): every rank executes exactly the code it executes under torchrun -- its own column shards,
its own leaf matrices, the row-sharded quotient / FRI kernels, the splicing of the query openings -- only the transport of
the collectives differs.  The NCCL + CUDA-IPC transport itself is exercised by bench.py --gpus N (parity_check)."""
import numpy as np
import pytest

from helpers import run_ranks

pytestmark = pytest.mark.gpu


def _single_gpu_proof(E, s):
    circ = E.Circuit.build(s)
    proof, _ = circ.prove(s["wires"], s["pi_hash"])
    circ.verify(s["pi_hash"], proof)
    return proof, np.array(circ.constants_sigmas.merkle_tree.cap)


@pytest.mark.parametrize("which,db,world", [("v1", 9, 1), ("v1", 9, 2), ("v2", 10, 4), ("v1", 11, 8), ("v2", 8, 8)])
def test_sharded_prover_equals_single_gpu(engine, which, db, world):
    E = engine
    s = E.synth_circuit(db, seed=40 + db) if which == "v1" else E.synth_circuit_v2(db, seed=40 + db)
    ref, cs_cap = _single_gpu_proof(E, s)

    def rank_fn(rank, dist):
        pr = E.ShardedProver(s["blob"], s["constants"], s["sigmas"], rank, world, dist=dist, use_peer=False)
        assert (np.array(pr.cs.cap) == cs_cap).all()
        proof, ms = pr.prove(s["wires"], s["pi_hash"])
        pr.verify(s["pi_hash"], proof)
        again, _ = pr.prove(s["wires"], s["pi_hash"])          # the handle is reusable
        assert (again == proof).all() and ms["total"] > 0
        pr.close()
        return proof

    for proof in run_ranks(world, rank_fn):
        assert proof.shape == ref.shape and (proof == ref).all()


def test_sharded_prover_limits(engine):
    E = engine
    s = E.synth_circuit(8, seed=3)
    with pytest.raises(E.EngineError, match="2\\^3 ranks|at most"):
        E.ShardedProver(s["blob"], s["constants"], s["sigmas"], 0, 16, use_peer=False)
    pr = E.ShardedProver(s["blob"], s["constants"], s["sigmas"], 0, 1, use_peer=False)
    with pytest.raises(E.EngineError, match="wire columns"):
        pr.prove(s["wires"][:100], s["pi_hash"])
    # the sharded handle has no local constants||sigmas commitment: the single-GPU entry points refuse it
    import ctypes as C
    from eth_lc_plonky2_b200 import _lib
    blob, n = C.POINTER(C.c_uint64)(), C.c_size_t(0)
    cols = [np.ascontiguousarray(c) for c in s["wires"]]
    ptrs = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
    pi = np.ascontiguousarray(s["pi_hash"])
    assert _lib.lib().eng_prove(pr._h, ptrs, pi.ctypes.data_as(C.c_void_p), C.byref(blob), C.byref(n), None) == E.ENG_ERR_STATE
    pr.close()


def test_splice_initial_openings_round_trip(engine):
    """The FriProof blob with the query openings spliced in by the host equals the one the engine writes itself."""
    E = engine
    rng = np.random.default_rng(1)
    from helpers import rand_field
    vals = [rand_field(rng, (w, 1 << 9)) for w in (5, 3)]
    eb = [E.PolynomialBatch.from_values(list(v), 3, False, 4) for v in vals]
    zeta = (123456789, 987654321)
    inst = E.FriInstanceInfo([(zeta, [(0, p) for p in range(5)] + [(1, p) for p in range(3)])])
    ch = E.Challenger(); ch.observe_elements(np.arange(5, dtype=np.uint64))
    proof = E.PolynomialBatch.prove_openings(inst, eb, ch, E.FriParams(9, 3, 4, 8, 6))
    # strip the initial openings, splice them back
    b = [int(x) for x in proof.blob]
    stripped, i = [], 0
    r = b[i]; i += 1
    for _ in range(r):
        i += 1 + b[i]
    f = b[i]; i += 1 + 2 * f + 1
    q = b[i]; i += 1
    stripped += b[:i]
    for _ in range(q):
        o = b[i]; i += 1
        for _ in range(o):
            i += 1 + b[i]
            i += 1 + 4 * b[i]
        stripped.append(0)
        s_ = b[i]; j = i + 1
        for _ in range(s_):
            j += 1 + 2 * b[j]
            j += 1 + 4 * b[j]
        stripped += b[i:j]; i = j
    back = E.splice_initial_openings(np.array(stripped, np.uint64), [[(leaf, path) for leaf, path in init] for init, _ in proof.query_round_proofs])
    assert (back == proof.blob).all()
