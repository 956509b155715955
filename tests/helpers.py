import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
P = 0xFFFFFFFF00000001


def golden():
    with open(os.path.join(HERE, "golden", "vectors.json")) as f:
        return json.load(f)


def hx(a):
    return " ".join("%016x" % int(x) for x in np.asarray(a).ravel())


def unhx(s):
    return np.array([int(x, 16) for x in s.split()], np.uint64)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).astype("<u8").tobytes()).hexdigest()


def structured(C, n):
    return np.arange(C, dtype=np.uint64)[:, None] * np.uint64(n) + np.arange(n, dtype=np.uint64)[None, :]


def bitrev(x, bits):
    return int(bin(x)[2:].zfill(bits)[::-1], 2) if bits else 0


def rand_field(rng, shape, noncanonical=False):
    """uniform u64; with noncanonical=True values in [p, 2^64) are kept (GoldilocksField allows them)."""
    a = rng.integers(0, 2**64 - 1, size=shape, dtype=np.uint64, endpoint=True)
    if not noncanonical:
        a = np.where(a >= np.uint64(P), a - np.uint64(P), a)
    return a


def pymul(a, b):
    return (int(a) * int(b)) % P


def poly_eval(coeffs, x):
    acc = 0
    for c in reversed([int(v) for v in coeffs]):
        acc = (acc * x + c) % P
    return acc
