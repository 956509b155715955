"""a9: the engine's host-side Challenger against the oracle's (runs on CPU: the transcript is host logic)."""
import numpy as np

from helpers import rand_field


def test_engine_challenger_matches_oracle(oracle):
    import eth_lc_plonky2_b200 as E
    rng = np.random.default_rng(8)
    a, b = E.Challenger(), oracle.Challenger()
    for step in range(40):
        k = int(rng.integers(0, 20))
        xs = rand_field(rng, k, noncanonical=True)
        if k:
            a.observe_elements(xs); b.observe(xs)
        m = int(rng.integers(0, 11))
        assert a.get_n_challenges(m) == b.get_n_challenges(m)
        assert (a.state() == b.state()).all()
    assert a.get_extension_challenge() == b.get_extension_challenge()
    c = E.Challenger()
    c.set_state(a.state())                       # state crosses the FFI (Rust Challenger <-> engine)
    assert c.get_n_challenges(9) == a.get_n_challenges(9)


def test_reduction_arity_bits(oracle):
    import eth_lc_plonky2_b200 as E
    for d in range(1, 25):
        assert E.reduction_arity_bits(d, 3, 4) == oracle.fri_arity_bits(d, 3, 4)
    assert E.reduction_arity_bits(22, 3, 4) == [4] * 5 and E.reduction_arity_bits(20, 3, 4) == [4] * 4
