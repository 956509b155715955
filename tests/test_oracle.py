"""CPU tests of the oracle (oracle/) against the golden vectors and algebraic properties.

The reference's own tests are prove -> verify round trips only (/root/reference/eth-lc-plonky2/src/unit_tests.rs:29-35)
and pin no prover output, so the pins are: plonky2's upstream Poseidon KATs (strong) and the SURVEY Appendix C vectors
(a second, independent restatement).  Everything else is property-checked.
"""
import numpy as np
import pytest

from helpers import P, bitrev, golden, hx, poly_eval, pymul, rand_field, sha, structured, unhx


def test_poseidon_upstream_kats(oracle):
    for inp, want in golden()["poseidon_kats"]:
        st = unhx(inp)
        assert hx(oracle.poseidon(st)) == want            # plonky2's production (fast partial rounds) form
        assert hx(oracle.poseidon(st, naive=True)) == want  # reference round function


def test_poseidon_noncanonical_inputs(oracle):
    rng = np.random.default_rng(7)
    st = rand_field(rng, (20, 12), noncanonical=True)
    st[0] = 2**64 - 1
    for s in st:
        canon = np.where(s >= np.uint64(P), s - np.uint64(P), s)
        assert (oracle.poseidon(s) == oracle.poseidon(canon)).all()
        assert (oracle.poseidon(s) == oracle.poseidon(s, naive=True)).all()


def test_sponge_vectors(oracle):
    g = golden()["sponge"]
    assert hx(oracle.hash_no_pad([1, 2, 3])) == g["hash_no_pad_1_2_3"]
    assert hx(oracle.hash_no_pad(np.arange(135))) == g["hash_no_pad_0_to_134"]
    h = oracle.hash_no_pad([1, 2, 3])
    assert hx(oracle.two_to_one(h, h)) == g["two_to_one_h_h"]
    # hash_or_noop: <= 4 elements are padded, not hashed; 5 are hashed
    assert hx(oracle.hash_or_noop([5, 6, 7])) == hx([5, 6, 7, 0])
    assert hx(oracle.hash_or_noop([1, 2, 3, 4, 5])) == hx(oracle.hash_no_pad([1, 2, 3, 4, 5]))
    # overwrite mode: a 9-element input = absorb 8, permute, overwrite lane 0 only, permute
    s = np.zeros(12, np.uint64); s[:8] = np.arange(1, 9)
    s = oracle.poseidon(s); s[0] = 9
    assert hx(oracle.hash_no_pad(np.arange(1, 10))) == hx(oracle.poseidon(s)[:4])


def test_structured_commit_vectors(oracle):
    for c in golden()["commits_structured"]:
        b = oracle.Batch.from_values(structured(c["C"], c["n"]), c["rate_bits"], c["cap_height"])
        assert hx(b.coeffs[1][:3]) == c["coeffs_1_0_3"]
        assert hx(b.leaves[1][:3]) == c["leaf_1_0_3"]
        assert hx(b.digests[0]) == c["digests_0"]
        assert hx(b.cap[0]) == c["cap_0"] and hx(b.cap[-1]) == c["cap_last"]
        assert sha(b.cap) == c["sha256_cap"] and sha(b.leaves) == c["sha256_leaves"]
        assert hx(b.leaves[-1][-1:]) == c["last_leaf_last_col"]


def test_generated_commit_pins(oracle):
    for g in golden()["oracle_generated"]:
        vals = oracle.splitmix_columns(g["C"], 1 << g["log_n"])
        b = oracle.Batch.from_values(vals, g["rate_bits"], g["cap_height"])
        assert sha(b.coeffs) == g["sha256_coeffs"] and sha(b.leaves) == g["sha256_leaves"]
        assert sha(b.digests) == g["sha256_digests"] and sha(b.cap) == g["sha256_cap"]


def test_field_constants(oracle):
    L = oracle.lib()
    g = L.orc_root_of_unity(32)
    assert g == 1753635133440165772 and pow(7, (P - 1) >> 32, P) == g
    for k, want in ((20, 3511170319078647661), (22, 5416168637041100469), (23, 16905767614792059275), (25, 5456943929260765144)):
        assert L.orc_root_of_unity(k) == want
    assert L.orc_gl_inv(7) == 2635249152773512046
    rng = np.random.default_rng(3)
    for a, b in rand_field(rng, (200, 2), noncanonical=True):
        assert L.orc_gl_mul(int(a), int(b)) == pymul(int(a) % P, int(b) % P)


@pytest.mark.parametrize("log_n", [0, 1, 3, 6, 10])
def test_fft_definition_and_inverse(oracle, log_n):
    rng = np.random.default_rng(log_n)
    n = 1 << log_n
    c = rand_field(rng, n)
    v = oracle.fft(c)
    w = oracle.lib().orc_root_of_unity(log_n)
    for i in ([0, 1, n - 1, n // 2] if n > 2 else range(n)):
        assert int(v[i]) == poly_eval(c, pow(w, i, P))   # values[i] = sum_k c_k w^{ik}, natural order
    assert (oracle.ifft(v) == c).all()


def test_lde_is_coset_evaluation(oracle):
    rng = np.random.default_rng(11)
    log_n, r = 5, 3
    c = rand_field(rng, 1 << log_n)
    v = oracle.lde(c, r)
    wl = oracle.lib().orc_root_of_unity(log_n + r)
    for j in (0, 1, 7, 100, (1 << (log_n + r)) - 1):
        assert int(v[j]) == poly_eval(c, 7 * pow(wl, j, P) % P)


@pytest.mark.parametrize("C,log_n,r,h", [(3, 2, 1, 0), (9, 3, 1, 1), (5, 4, 2, 6), (135, 4, 3, 4), (2, 5, 0, 5)])
def test_batch_semantics(oracle, C, log_n, r, h):
    """leaves[k] = evaluations at 7*w_L^{bitrev(k)}; every Merkle path verifies against the cap."""
    rng = np.random.default_rng(C * 100 + log_n)
    vals = rand_field(rng, (C, 1 << log_n))
    b = oracle.Batch.from_values(vals, r, h)
    log_l = log_n + r
    wl = oracle.lib().orc_root_of_unity(log_l)
    for k in (0, 1, (1 << log_l) - 1, 5 % (1 << log_l)):
        x = 7 * pow(wl, bitrev(k, log_l), P) % P
        for c in (0, C - 1):
            assert int(b.leaves[k][c]) == poly_eval(b.coeffs[c], x)
    wn = oracle.lib().orc_root_of_unity(log_n)
    for c in (0, C - 1):
        for i in (0, (1 << log_n) - 1):
            assert poly_eval(b.coeffs[c], pow(wn, i, P)) == int(vals[c][i])   # from_values interpolates
    assert b.digests.shape[0] == 2 * ((1 << log_l) - (1 << h))
    for k in range(0, 1 << log_l, max(1, (1 << log_l) // 9)):
        sib = b.prove(k)
        assert sib.shape[0] == log_l - h
        assert oracle.merkle_verify(b.leaves[k], k, b.cap, sib)
        if sib.shape[0]:
            bad = sib.copy(); bad[0, 0] ^= np.uint64(1)
            assert not oracle.merkle_verify(b.leaves[k], k, b.cap, bad)


def test_merkle_preconditions(oracle):
    leaves = np.arange(8 * 5, dtype=np.uint64).reshape(8, 5)
    oracle.MerkleTree(leaves, 3)                    # cap == leaf digests
    with pytest.raises(ValueError):
        oracle.MerkleTree(leaves, 4)                # cap_height > log2(leaves): plonky2 panics
    t = oracle.MerkleTree(leaves, 3)
    for i in range(8):
        assert hx(t.cap[i]) == hx(oracle.hash_or_noop(leaves[i]))


def test_challenger_duplex(oracle):
    ch = oracle.Challenger()
    ch.observe([1, 2, 3])
    c0 = ch.get_challenge()
    s = np.zeros(12, np.uint64); s[:3] = [1, 2, 3]
    out = oracle.poseidon(s)
    assert c0 == int(out[7])                        # challenges pop from the END of the rate part
    assert ch.get_challenge() == int(out[6])
    ch.observe([9])                                 # observing clears the output buffer
    s2 = out.copy(); s2[0] = 9
    assert ch.get_challenge() == int(oracle.poseidon(s2)[7])
    ch2 = oracle.Challenger()
    ch2.observe(np.arange(8))                       # 8 inputs trigger a duplex immediately
    s3 = np.zeros(12, np.uint64); s3[:8] = np.arange(8)
    assert ch2.get_challenge() == int(oracle.poseidon(s3)[7])
