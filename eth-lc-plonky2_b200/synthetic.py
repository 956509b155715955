"""Synthetic witness columns for benchmarks and tests: values[c][i] = SplitMix64 seeded with
`seed ^ c`, step i, reduced mod p.  Input generation only -- no prover arithmetic."""
import numpy as np

P = 0xFFFFFFFF00000001


def splitmix_columns(num_cols, n, seed=0x9E3779B97F4A7C15, first_col=0):
    """Columns first_col .. first_col + num_cols - 1 of the synthetic witness."""
    out = np.empty((num_cols, n), np.uint64)
    steps = np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    with np.errstate(over="ignore"):
        for c in range(num_cols):
            z = np.uint64((seed ^ (first_col + c)) & 0xFFFFFFFFFFFFFFFF) + steps
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
            out[c] = np.where(z >= np.uint64(P), z - np.uint64(P), z)
    return out
