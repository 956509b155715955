"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY (see oracle/gl.h).

A C++ restatement of plonky2 0.1.4's commit path.  Builds itself with `make`
on first use if the shared object is missing.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
P = 0xFFFFFFFF00000001

_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


def build(force=False):
    import fcntl
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".h", ".cpp"))]
    # several test processes (gloo workers) may get here at once: one builds, the others wait and then find it fresh
    with open(os.path.join(_HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
            subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_poseidon_permute.argtypes = [_u64p, C.c_int]
        L.orc_hash_n.argtypes = [_u64p, C.c_size_t, _u64p, C.c_int]
        L.orc_compress.argtypes = [_u64p, _u64p, _u64p]
        for f in (L.orc_gl_mul, L.orc_gl_pow):
            f.restype = C.c_uint64; f.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_gl_inv.restype = C.c_uint64; L.orc_gl_inv.argtypes = [C.c_uint64]
        L.orc_root_of_unity.restype = C.c_uint64; L.orc_root_of_unity.argtypes = [C.c_int]
        L.orc_fft.argtypes = [_u64p, C.c_int]
        L.orc_ifft.argtypes = [_u64p, C.c_int]
        L.orc_lde.argtypes = [_u64p, C.c_int, C.c_int, C.c_uint64, _u64p]
        for f in (L.orc_batch_from_values_c, L.orc_batch_from_coeffs_c):
            f.restype = C.c_void_p; f.argtypes = [_u64p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_batch_free.argtypes = [C.c_void_p]
        for f in (L.orc_batch_coeffs, L.orc_batch_leaves, L.orc_batch_digests, L.orc_batch_cap):
            f.restype = C.POINTER(C.c_uint64); f.argtypes = [C.c_void_p]
        L.orc_batch_num_digests.restype = C.c_size_t; L.orc_batch_num_digests.argtypes = [C.c_void_p]
        L.orc_batch_prove.restype = C.c_int; L.orc_batch_prove.argtypes = [C.c_void_p, C.c_size_t, _u64p]
        L.orc_batch_times.argtypes = [C.c_void_p, np.ctypeslib.ndpointer(dtype=np.float64)]
        L.orc_merkle_new.restype = C.c_void_p; L.orc_merkle_new.argtypes = [_u64p, C.c_size_t, C.c_size_t, C.c_int]
        L.orc_merkle_verify_c.restype = C.c_int
        L.orc_merkle_verify_c.argtypes = [_u64p, C.c_size_t, C.c_size_t, _u64p, _u64p, C.c_int]
        L.orc_challenger_new.restype = C.c_void_p
        L.orc_challenger_free.argtypes = [C.c_void_p]
        L.orc_challenger_observe.argtypes = [C.c_void_p, _u64p, C.c_size_t]
        L.orc_challenger_get.restype = C.c_uint64; L.orc_challenger_get.argtypes = [C.c_void_p]
        L.orc_batch_eval_c.argtypes = [C.c_void_p, _u64p, _u64p]
        L.orc_fri_arity_bits_c.restype = C.c_int
        L.orc_fri_arity_bits_c.argtypes = [C.c_int] * 5 + [C.POINTER(C.c_int)]
        L.orc_fri_prove_c.restype = C.c_void_p
        L.orc_fri_prove_c.argtypes = [_u64p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        L.orc_blob_len.restype = C.c_size_t; L.orc_blob_len.argtypes = [C.c_void_p]
        L.orc_blob_data.restype = C.POINTER(C.c_uint64); L.orc_blob_data.argtypes = [C.c_void_p]
        L.orc_blob_free.argtypes = [C.c_void_p]
        L.orc_fri_verify_c.restype = C.c_int
        L.orc_fri_verify_c.argtypes = [_u64p, _u64p, _u64p, C.POINTER(C.c_int), C.c_int, _u64p, C.c_size_t, C.c_void_p, C.POINTER(C.c_int)]
        L.orc_challenger_state.argtypes = [C.c_void_p, _u64p]
        L.orc_circuit_new.restype = C.c_void_p; L.orc_circuit_new.argtypes = [_u64p]
        L.orc_circuit_free.argtypes = [C.c_void_p]
        L.orc_partial_products_c.argtypes = [C.c_void_p, _u64p, _u64p, _u64p, _u64p, _u64p]
        L.orc_quotient_c.restype = C.c_int
        L.orc_quotient_c.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _u64p, _u64p, _u64p, _u64p, _u64p]
        L.orc_prove_c.restype = C.c_void_p; L.orc_prove_c.argtypes = [C.c_void_p, C.c_void_p, _u64p, _u64p, _u64p]
        L.orc_verify_c.restype = C.c_int; L.orc_verify_c.argtypes = [C.c_void_p, _u64p, _u64p, _u64p, C.c_size_t]
        _lib = L
    return _lib


def _a(x):
    return np.ascontiguousarray(x, dtype=np.uint64)


def poseidon(state, naive=False):
    s = _a(state).copy()
    assert s.shape == (12,)
    lib().orc_poseidon_permute(s, int(naive))
    return s


def hash_no_pad(xs):
    out = np.zeros(4, np.uint64); xs = _a(xs)
    lib().orc_hash_n(xs, xs.size, out, 0)
    return out


def hash_or_noop(xs):
    out = np.zeros(4, np.uint64); xs = _a(xs)
    lib().orc_hash_n(xs, xs.size, out, 1)
    return out


def two_to_one(l, r):
    out = np.zeros(4, np.uint64)
    lib().orc_compress(_a(l), _a(r), out)
    return out


def fft(a):
    a = _a(a).copy(); lib().orc_fft(a, int(a.size).bit_length() - 1); return a


def ifft(a):
    a = _a(a).copy(); lib().orc_ifft(a, int(a.size).bit_length() - 1); return a


def lde(coeffs, rate_bits, shift=7):
    coeffs = _a(coeffs); log_n = int(coeffs.size).bit_length() - 1
    out = np.zeros(coeffs.size << rate_bits, np.uint64)
    lib().orc_lde(coeffs, log_n, rate_bits, shift, out)
    return out


class Batch:
    """PolynomialBatch as computed by the oracle: coeffs [C][n], leaves [L][C], digests, cap."""

    def __init__(self, handle, C_, log_n, rate_bits, cap_height):
        if not handle:
            raise ValueError("oracle: MerkleTree::new precondition violated (cap_height > log2(leaves))")
        self.h, self.C, self.log_n, self.rate_bits, self.cap_height = handle, C_, log_n, rate_bits, cap_height
        self.n = 1 << log_n
        self.L = self.n << rate_bits

    @classmethod
    def from_values(cls, values, rate_bits, cap_height):
        values = _a(values); C_, n = values.shape
        log_n = n.bit_length() - 1
        return cls(lib().orc_batch_from_values_c(values, C_, log_n, rate_bits, cap_height), C_, log_n, rate_bits, cap_height)

    @classmethod
    def from_coeffs(cls, coeffs, rate_bits, cap_height):
        coeffs = _a(coeffs); C_, n = coeffs.shape
        log_n = n.bit_length() - 1
        return cls(lib().orc_batch_from_coeffs_c(coeffs, C_, log_n, rate_bits, cap_height), C_, log_n, rate_bits, cap_height)

    def _view(self, ptr, shape):
        return np.ctypeslib.as_array(ptr, shape=shape)

    @property
    def coeffs(self):
        return self._view(lib().orc_batch_coeffs(self.h), (self.C, self.n))

    @property
    def leaves(self):
        return self._view(lib().orc_batch_leaves(self.h), (self.L, self.C))

    @property
    def digests(self):
        nd = lib().orc_batch_num_digests(self.h)
        if nd == 0:
            return np.zeros((0, 4), np.uint64)
        return self._view(lib().orc_batch_digests(self.h), (nd, 4))

    @property
    def cap(self):
        return self._view(lib().orc_batch_cap(self.h), (1 << self.cap_height, 4))

    def get(self, leaf_index):
        return self.leaves[leaf_index]

    def prove(self, leaf_index):
        sib = np.zeros((64, 4), np.uint64)
        k = lib().orc_batch_prove(self.h, leaf_index, sib)
        return sib[:k].copy()

    def times(self):
        t = np.zeros(4); lib().orc_batch_times(self.h, t)
        return dict(ifft=t[0], lde=t[1], transpose=t[2], tree=t[3])

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_batch_free(self.h); self.h = None


class MerkleTree(Batch):
    def __init__(self, leaves, cap_height):
        leaves = _a(leaves); L, w = leaves.shape
        h = lib().orc_merkle_new(leaves, L, w, cap_height)
        if not h:
            raise ValueError("oracle: MerkleTree::new precondition violated (cap_height > log2(leaves))")
        self.h, self.C, self.L, self.cap_height = h, w, L, cap_height
        self.n, self.log_n, self.rate_bits = L, L.bit_length() - 1, 0


def merkle_verify(leaf, leaf_index, cap, siblings):
    leaf = _a(leaf); siblings = _a(siblings).reshape(-1, 4)
    return bool(lib().orc_merkle_verify_c(leaf, leaf.size, leaf_index, _a(cap), siblings, siblings.shape[0]))


class Challenger:
    def __init__(self):
        self.h = lib().orc_challenger_new()

    def observe(self, xs):
        xs = _a(np.atleast_1d(xs)).ravel(); lib().orc_challenger_observe(self.h, xs, xs.size)

    def get_challenge(self):
        return int(lib().orc_challenger_get(self.h))

    def get_n_challenges(self, n):
        return [self.get_challenge() for _ in range(n)]

    def get_extension_challenge(self):
        return (self.get_challenge(), self.get_challenge())

    def state(self):
        out = np.zeros(30, np.uint64)
        lib().orc_challenger_state(self.h, out)
        return out

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_challenger_free(self.h); self.h = None


def poly_eval_base(coeffs, x):
    """PolynomialCoeffs::eval at a base-field point (Horner), canonical result."""
    c = np.ascontiguousarray(coeffs, dtype=np.uint64)
    f = lib().orc_poly_eval
    f.argtypes = [_u64p, C.c_uint64, C.c_uint64]
    f.restype = C.c_uint64
    return int(f(c, c.shape[0], int(x)))


def batch_eval(batch, z):
    """eval_commitment(z, batch): every polynomial of the batch at z in F_p^2 -> [C][2]."""
    out = np.zeros((batch.C, 2), np.uint64)
    lib().orc_batch_eval_c(batch.h, _a(z), out)
    return out


def fri_arity_bits(degree_bits, rate_bits=3, cap_height=4, arity_bits=4, final_poly_bits=5):
    out = (C.c_int * 32)()
    k = lib().orc_fri_arity_bits_c(degree_bits, rate_bits, cap_height, arity_bits, final_poly_bits, out)
    return list(out[:k])


def fri_instance_blob(batches):
    """batches: [(point (a, b), [(oracle, poly), ...]), ...] -> flat u64 description shared with the engine."""
    o = [len(batches)]
    for point, polys in batches:
        o += [int(point[0]), int(point[1]), len(polys)] + [(int(a) << 32) | int(b) for a, b in polys]
    return np.array(o, np.uint64)


def fri_params_array(degree_bits, rate_bits, cap_height, pow_bits, num_query_rounds, arity_bits):
    v = [degree_bits, rate_bits, cap_height, pow_bits, num_query_rounds, len(arity_bits)] + list(arity_bits)
    return (C.c_int * len(v))(*v)


def fri_prove(instance_blob, oracles, challenger, params):
    hs = (C.c_void_p * len(oracles))(*[o.h for o in oracles])
    b = lib().orc_fri_prove_c(_a(instance_blob), hs, len(oracles), challenger.h, params)
    n = lib().orc_blob_len(b)
    out = np.ctypeslib.as_array(lib().orc_blob_data(b), shape=(n,)).copy()
    lib().orc_blob_free(b)
    return out


def fri_verify(instance_blob, openings, caps, leaf_lens, proof_blob, challenger, params):
    """0 = verifies; otherwise the code of the failed check (see oracle/fri.h::orc_fri_verify)."""
    ll = (C.c_int * len(leaf_lens))(*leaf_lens)
    proof_blob = _a(proof_blob)
    return lib().orc_fri_verify_c(_a(instance_blob), _a(openings).ravel(), _a(caps).ravel(), ll, len(leaf_lens), proof_blob,
                                  proof_blob.size, challenger.h, params)


class Circuit:
    """Circuit description for the plonk rows (oracle/plonk.h); `blob` layout documented at orc_circuit_new."""

    def __init__(self, blob):
        self.blob = _a(blob)
        self.h = lib().orc_circuit_new(self.blob)
        b = [int(x) for x in self.blob]
        if b[0] == 0x32424B4C50:     # version-2 description: header after [magic, length]
            b = b[2:]
        (self.degree_bits, self.num_wires, self.num_routed, self.num_gate_constants, self.num_selectors, self.num_challenges,
         self.quotient_degree_factor, self.rate_bits, self.cap_height) = b[:9]
        self.num_pp = (self.num_routed + self.quotient_degree_factor - 1) // self.quotient_degree_factor - 1

    def partial_products(self, wires, sigmas, betas, gammas):
        n = 1 << self.degree_bits
        out = np.zeros((self.num_challenges * (1 + self.num_pp), n), np.uint64)
        lib().orc_partial_products_c(self.h, _a(wires), _a(sigmas), _a(betas), _a(gammas), out)
        return out

    def quotient(self, cs, wires, zs_pp, pi_hash, betas, gammas, alphas):
        n = 1 << self.degree_bits
        out = np.zeros((self.num_challenges * self.quotient_degree_factor, n), np.uint64)
        rc = lib().orc_quotient_c(self.h, cs.h, wires.h, zs_pp.h, _a(pi_hash), _a(betas), _a(gammas), _a(alphas), out)
        if rc:
            raise ValueError("oracle: quotient_degree_bits != rate_bits is not restated")
        return out

    def prove(self, cs, wire_values, sigma_values, pi_hash):
        b = lib().orc_prove_c(self.h, cs.h, _a(wire_values), _a(sigma_values), _a(pi_hash))
        if not b:
            raise ValueError("oracle: prove failed")
        n = lib().orc_blob_len(b)
        out = np.ctypeslib.as_array(lib().orc_blob_data(b), shape=(n,)).copy()
        lib().orc_blob_free(b)
        return out

    def verify(self, cs_cap, pi_hash, proof_blob):
        proof_blob = _a(proof_blob)
        return lib().orc_verify_c(self.h, _a(cs_cap).ravel(), _a(pi_hash), proof_blob, proof_blob.size)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_circuit_free(self.h); self.h = None


def splitmix_columns(C_, n, seed=0x9E3779B97F4A7C15):
    """Synthetic witness of SURVEY.md 8(d): values[c][i] = SplitMix64(seed ^ c) step i, reduced mod p."""
    M = np.uint64(0xFFFFFFFFFFFFFFFF)
    out = np.empty((C_, n), np.uint64)
    steps = (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
    with np.errstate(over="ignore"):
        for c in range(C_):
            z = np.uint64(seed ^ c) + steps
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
            out[c] = np.where(z >= np.uint64(P), z - np.uint64(P), z)
    return out
