"""Host-side mirror of the plonky2 operator surface that the engine replaces.

Names, argument meaning and failure behaviour follow plonky2 0.1.4 (dep pinned at
/root/reference/Cargo.lock:2347-2350; the reference reaches these through builder.build() and data.prove(),
/root/reference/eth-lc-plonky2/src/main.rs:227,230):

    PolynomialBatch::from_values / from_coeffs / get_lde_values      [plonky2:fri/oracle.rs]
    MerkleTree::new / cap / get / prove, digests layout              [plonky2:hash/merkle_tree.rs]
    PoseidonHash::hash_no_pad / hash_or_noop / two_to_one, poseidon  [plonky2:hash/poseidon.rs, hashing.rs]

Everything is computed by the CUDA engine through the C ABI (include/plonky2_b200.h); batches stay resident
on the device behind handles and the accessors below are gathers.  A violated plonky2 precondition (an
assert!/expect panic in Rust) raises EngineError with status ENG_ERR_INVALID.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import EngineError, check, host_u64, ptr

SALT_SIZE = 4
P = 0xFFFFFFFF00000001


def _is_device_tensor(x):
    return hasattr(x, "is_cuda") and x.is_cuda


def poseidon(states):
    """Poseidon::poseidon on one state (12,) or many (k, 12)."""
    s = host_u64(states)
    single = s.ndim == 1
    s = s.reshape(-1, 12)
    out = np.empty_like(s)
    check(_lib.lib().eng_poseidon_permute(ptr(s), ptr(out), s.shape[0]))
    return out[0] if single else out


class PoseidonHash:
    """plonky2::hash::poseidon::PoseidonHash (Hasher for PoseidonGoldilocksConfig)."""

    @staticmethod
    def _hash(inputs, or_noop):
        a = host_u64(inputs)
        single = a.ndim == 1
        a = a.reshape(1, -1) if single else a
        out = np.empty((a.shape[0], 4), np.uint64)
        check(_lib.lib().eng_hash_n(ptr(a), a.shape[1], a.shape[0], int(or_noop), ptr(out)))
        return out[0] if single else out

    @staticmethod
    def hash_no_pad(inputs):
        return PoseidonHash._hash(inputs, False)

    @staticmethod
    def hash_or_noop(inputs):
        return PoseidonHash._hash(inputs, True)

    @staticmethod
    def two_to_one(left, right):
        l, r = host_u64(left).reshape(-1, 4), host_u64(right).reshape(-1, 4)
        pairs = np.ascontiguousarray(np.concatenate([l, r], axis=1))
        out = np.empty((pairs.shape[0], 4), np.uint64)
        check(_lib.lib().eng_two_to_one(ptr(pairs), pairs.shape[0], ptr(out)))
        return out[0] if np.ndim(left) == 1 else out


class _Handle:
    """Owner of an eng_batch* (Rust: Drop -> eng_batch_free)."""

    def __init__(self, handle):
        self._h = handle
        info = _lib.BatchInfo()
        check(_lib.lib().eng_batch_info(self._h, C.byref(info)))
        self.info = info

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().eng_batch_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MerkleTree:
    """plonky2::hash::merkle_tree::MerkleTree<F, PoseidonHash>."""

    def __init__(self, handle_owner):
        self._o = handle_owner

    @classmethod
    def new(cls, leaves, cap_height):
        """MerkleTree::new(leaves: Vec<Vec<F>>, cap_height).  `leaves` is [num_leaves][leaf_len] (numpy, host)."""
        a = host_u64(leaves)
        if a.ndim != 2:
            raise EngineError(_lib.ENG_ERR_INVALID, "leaves must be [num_leaves][leaf_len]")
        h = C.c_void_p()
        check(_lib.lib().eng_merkle_new(ptr(a), a.shape[0], a.shape[1], cap_height, C.byref(h)))
        return cls(_Handle(h))

    @property
    def cap_height(self):
        return self._o.info.cap_height

    @property
    def num_leaves(self):
        return self._o.info.num_leaves

    @property
    def cap(self):
        """MerkleCap: [2^cap_height][4]."""
        out = np.empty((1 << self._o.info.cap_height, 4), np.uint64)
        check(_lib.lib().eng_batch_cap(self._o._h, ptr(out)))
        return out

    @property
    def digests(self):
        """The `digests` vector in plonky2's interleaved order: [2*(L - 2^h)][4]."""
        out = np.empty((self._o.info.num_digests, 4), np.uint64)
        check(_lib.lib().eng_batch_digests(self._o._h, ptr(out)))
        return out

    def get(self, i):
        """MerkleTree::get(i) -> &leaves[i]."""
        return self.leaves(i, 1)[0]

    def leaves(self, first=0, count=None):
        """leaves[first : first+count], row-major."""
        if count is None:
            count = self._o.info.num_leaves - first
        out = np.empty((count, self._o.info.leaf_len), np.uint64)
        check(_lib.lib().eng_batch_leaves(self._o._h, first, count, ptr(out)))
        return out

    def prove(self, leaf_index):
        """MerkleTree::prove(leaf_index).siblings, bottom-up: [log2(L) - cap_height][4]."""
        n = C.c_uint32(0)
        out = np.empty((64, 4), np.uint64)
        check(_lib.lib().eng_batch_merkle_path(self._o._h, leaf_index, ptr(out), C.byref(n)))
        return out[: n.value].copy()


class PolynomialBatch:
    """plonky2::fri::oracle::PolynomialBatch<F, C, D> resident on the device."""

    def __init__(self, handle_owner):
        self._o = handle_owner
        self.merkle_tree = MerkleTree(handle_owner)

    @staticmethod
    def _make(data, rate_bits, blinding, cap_height, is_values, blinding_seed):
        lib = _lib.lib()
        h = C.c_void_p()
        if _is_device_tensor(data):
            if data.dim() != 2 or data.element_size() != 8 or not data.is_contiguous():
                raise EngineError(_lib.ENG_ERR_INVALID, "device input must be a contiguous [num_polys][n] 64-bit tensor")
            num_polys, n = data.shape
            fn = lib.eng_batch_from_values_dev if is_values else lib.eng_batch_from_coeffs_dev
            arg = C.c_void_p(data.data_ptr())
            keep = None
        else:
            cols = [host_u64(c) for c in data]
            if not cols:
                raise EngineError(_lib.ENG_ERR_INVALID, "PolynomialBatch needs at least one polynomial")
            n = cols[0].shape[0]
            if any(c.ndim != 1 or c.shape[0] != n for c in cols):
                # plonky2: "All polynomials must have the same length"
                raise EngineError(_lib.ENG_ERR_INVALID, "all polynomials must have the same length")
            num_polys = len(cols)
            fn = lib.eng_batch_from_values if is_values else lib.eng_batch_from_coeffs
            keep = (C.c_void_p * num_polys)(*[c.ctypes.data for c in cols])
            arg = keep
        if n == 0 or n & (n - 1):
            raise EngineError(_lib.ENG_ERR_INVALID, "polynomial length %d is not a power of two (log2_strict)" % n)
        log_n = n.bit_length() - 1
        check(fn(arg, num_polys, log_n, rate_bits, int(bool(blinding)), blinding_seed, cap_height, C.byref(h)))
        return PolynomialBatch(_Handle(h))

    @classmethod
    def from_values(cls, values, rate_bits, blinding, cap_height, timing=None, fft_root_table=None, blinding_seed=0):
        """PolynomialBatch::from_values(values: Vec<PolynomialValues<F>>, rate_bits, blinding, cap_height, ..).
        blinding_seed = 0 (the default): salts keyed from the OS RNG, as plonky2's OsRng; non-zero: reproducible test stream."""
        return cls._make(values, rate_bits, blinding, cap_height, True, blinding_seed)

    @classmethod
    def from_coeffs(cls, polynomials, rate_bits, blinding, cap_height, timing=None, fft_root_table=None, blinding_seed=0):
        """PolynomialBatch::from_coeffs(polynomials: Vec<PolynomialCoeffs<F>>, rate_bits, blinding, cap_height, ..)."""
        return cls._make(polynomials, rate_bits, blinding, cap_height, False, blinding_seed)

    # ---- fields ----
    @property
    def degree_log(self):
        return self._o.info.degree_log

    @property
    def rate_bits(self):
        return self._o.info.rate_bits

    @property
    def blinding(self):
        return bool(self._o.info.blinding)

    @property
    def num_polys(self):
        return self._o.info.num_polys

    @property
    def polynomials(self):
        """Vec<PolynomialCoeffs<F>> as [num_polys][n] (device -> host copy)."""
        n = 1 << self._o.info.degree_log
        out = np.empty((self._o.info.num_polys, n), np.uint64)
        for c in range(self._o.info.num_polys):
            check(_lib.lib().eng_batch_coeffs(self._o._h, c, ptr(out[c])))
        return out

    def get_lde_values(self, index, step):
        """leaves[bitrev(index*step)] without the salt."""
        out = np.empty(self._o.info.num_polys, np.uint64)
        check(_lib.lib().eng_batch_lde_values(self._o._h, index, step, ptr(out)))
        return out

    def stage_ms(self):
        """Device time per stage, keyed by plonky2's timed! labels."""
        t = (C.c_float * 6)()
        check(_lib.lib().eng_batch_stage_ms(self._o._h, t))
        return {"IFFT": t[0], "FFT + blinding": t[1], "transpose LDEs": t[2], "build Merkle tree (leaves)": t[3],
                "build Merkle tree (digest levels)": t[4], "host to device": t[5]}

    def device_ptrs(self):
        p = [C.c_void_p() for _ in range(4)]
        check(_lib.lib().eng_batch_device_ptrs(self._o._h, *[C.byref(x) for x in p]))
        return {"lde": p[0].value, "coeffs": p[1].value, "digests": p[2].value, "cap": p[3].value}

    def close(self):
        self._o.close()


# ---------------------------------------------------------------- a9: Challenger
class Challenger:
    """plonky2::iop::challenger::Challenger<F, PoseidonHash> (host-side duplex transcript inside the engine library)."""

    def __init__(self):
        self._h = C.c_void_p()
        check(_lib.load().eng_challenger_new(C.byref(self._h)))

    def observe_elements(self, elements):
        a = host_u64(np.atleast_1d(elements)).ravel()
        check(_lib.load().eng_challenger_observe(self._h, ptr(a), a.size))

    observe_element = observe_elements
    observe_hash = observe_elements               # HashOut: 4 elements
    observe_cap = observe_elements                # MerkleCap: 2^h hashes, in order
    observe_extension_element = observe_elements  # to_basefield_array(): [a0, a1]
    observe_extension_elements = observe_elements

    def get_n_challenges(self, n):
        out = np.empty(n, np.uint64)
        check(_lib.load().eng_challenger_get_challenges(self._h, ptr(out), n))
        return [int(x) for x in out]

    def get_challenge(self):
        return self.get_n_challenges(1)[0]

    def get_extension_challenge(self):
        return tuple(self.get_n_challenges(2))

    def get_hash(self):
        return np.array(self.get_n_challenges(4), np.uint64)

    def state(self):
        out = np.empty(30, np.uint64)
        check(_lib.load().eng_challenger_get_state(self._h, ptr(out)))
        return out

    def set_state(self, st):
        st = host_u64(st)
        check(_lib.load().eng_challenger_set_state(self._h, ptr(st)))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.load().eng_challenger_free(self._h)
                self._h = None
        except Exception:
            pass


# ---------------------------------------------------------------- a7 / a8: openings and FRI
def reduction_arity_bits(degree_bits, rate_bits, cap_height, arity_bits=4, final_poly_bits=5):
    """FriReductionStrategy::ConstantArityBits(arity_bits, final_poly_bits).reduction_arity_bits(..)."""
    out = []
    while degree_bits > final_poly_bits and degree_bits + rate_bits - arity_bits >= cap_height:
        out.append(arity_bits)
        degree_bits -= arity_bits
    return out


class FriParams:
    """plonky2::fri::FriParams (+ the FriConfig fields the prover reads)."""

    def __init__(self, degree_bits, rate_bits=3, cap_height=4, proof_of_work_bits=16, num_query_rounds=28, arity_bits=None):
        self.degree_bits, self.rate_bits, self.cap_height = degree_bits, rate_bits, cap_height
        self.proof_of_work_bits, self.num_query_rounds = proof_of_work_bits, num_query_rounds
        self.reduction_arity_bits = reduction_arity_bits(degree_bits, rate_bits, cap_height) if arity_bits is None else list(arity_bits)

    def as_array(self):
        v = [self.degree_bits, self.rate_bits, self.cap_height, self.proof_of_work_bits, self.num_query_rounds,
             len(self.reduction_arity_bits)] + self.reduction_arity_bits
        return (C.c_int32 * len(v))(*v)


class FriInstanceInfo:
    """plonky2::fri::structure::FriInstanceInfo: batches = [(point (a, b), [(oracle_index, polynomial_index), ...]), ...]."""

    def __init__(self, batches):
        self.batches = batches

    def blob(self):
        o = [len(self.batches)]
        for point, polys in self.batches:
            o += [int(point[0]), int(point[1]), len(polys)] + [(int(a) << 32) | int(b) for a, b in polys]
        return np.array(o, np.uint64)


class FriProof:
    """plonky2::fri::proof::FriProof as the engine's flat blob (layout in include/plonky2_b200.h), parsed lazily."""

    def __init__(self, blob):
        self.blob = blob
        i = 0
        b = [int(x) for x in blob]
        r = b[i]; i += 1
        self.commit_phase_merkle_caps = []
        for _ in range(r):
            n = b[i]; i += 1
            self.commit_phase_merkle_caps.append(np.array(b[i:i + n], np.uint64).reshape(-1, 4)); i += n
        f = b[i]; i += 1
        self.final_poly = np.array(b[i:i + 2 * f], np.uint64).reshape(-1, 2); i += 2 * f
        self.pow_witness = b[i]; i += 1
        q = b[i]; i += 1
        self.query_round_proofs = []
        for _ in range(q):
            o = b[i]; i += 1
            initial = []
            for _ in range(o):
                n = b[i]; i += 1
                leaf = np.array(b[i:i + n], np.uint64); i += n
                n = b[i]; i += 1
                path = np.array(b[i:i + 4 * n], np.uint64).reshape(-1, 4); i += 4 * n
                initial.append((leaf, path))
            s = b[i]; i += 1
            steps = []
            for _ in range(s):
                n = b[i]; i += 1
                evals = np.array(b[i:i + 2 * n], np.uint64).reshape(-1, 2); i += 2 * n
                n = b[i]; i += 1
                path = np.array(b[i:i + 4 * n], np.uint64).reshape(-1, 4); i += 4 * n
                steps.append((evals, path))
            self.query_round_proofs.append((initial, steps))
        assert i == len(b)


def _eval_batch(self, z):
    """eval_commitment(z, self) of OpeningSet::new: every polynomial at z = (a, b) in F_p^2 -> [num_polys][2]."""
    zz = host_u64([int(z[0]), int(z[1])])
    out = np.empty((self._o.info.num_polys, 2), np.uint64)
    check(_lib.lib().eng_batch_eval_ext(self._o._h, ptr(zz), ptr(out)))
    return out


def _prove_openings(instance, oracles, challenger, fri_params, timing=None):
    """PolynomialBatch::prove_openings(instance, oracles, challenger, fri_params, timing) -> FriProof."""
    hs = (C.c_void_p * len(oracles))(*[o._o._h for o in oracles])
    inst = instance.blob()
    blob = C.POINTER(C.c_uint64)()
    n = C.c_size_t(0)
    check(_lib.lib().eng_fri_prove_openings(ptr(inst), hs, len(oracles), challenger._h, fri_params.as_array(), C.byref(blob), C.byref(n)))
    out = np.ctypeslib.as_array(blob, shape=(n.value,)).copy()
    _lib.lib().eng_blob_free(blob)
    return FriProof(out)


PolynomialBatch.eval = _eval_batch
PolynomialBatch.prove_openings = staticmethod(_prove_openings)


# ---------------------------------------------------------------- a5 / a6 / prove / verify
GATE_KINDS = ("Noop", "Constant", "PublicInput", "Arithmetic", "Poseidon", "BaseSum", "ArithmeticExtension", "MulExtension", "Reducing",
              "ReducingExtension", "RandomAccess", "Exponentiation", "PoseidonMds", "U32Arithmetic", "U32AddMany", "U32Subtraction",
              "U32RangeCheck", "Comparison", "CosetInterpolation", "U32Interleave", "UninterleaveToU32", "UninterleaveToB32")
BLOB_V2_MAGIC = 0x32424B4C50
ALL_GATES = (1 << len(GATE_KINDS)) - 1


def _take_blob(ptr, n):
    out = np.ctypeslib.as_array(ptr, shape=(n.value,)).copy()
    _lib.load().eng_blob_free(ptr)
    return out


def synth_circuit(degree_bits, seed=1):
    """Synthetic circuit over the five core gates with a satisfying witness (eng_synth_circuit; host code, no GPU).
    Returns dict(blob, constants [4][n], sigmas [80][n], wires [135][n], pi_hash [4])."""
    n = 1 << degree_bits
    out = dict(constants=np.zeros((4, n), np.uint64), sigmas=np.zeros((80, n), np.uint64), wires=np.zeros((135, n), np.uint64),
               pi_hash=np.zeros(4, np.uint64), blob=np.zeros(36, np.uint64))
    check(_lib.load().eng_synth_circuit(degree_bits, seed, ptr(out["constants"]), ptr(out["sigmas"]), ptr(out["wires"]),
                                        ptr(out["pi_hash"]), ptr(out["blob"])))
    return out


def synth_circuit_v2(degree_bits, seed=1, kinds_mask=ALL_GATES):
    """Synthetic circuit over the gates of `kinds_mask` (bit k = GATE_KINDS[k]; default: all 22) with a satisfying witness,
    selector groups formed by plonky2's rule, version-2 description with the gates' bytecode (eng_synth_circuit_v2; host code)."""
    n = 1 << degree_bits
    consts = np.zeros((8, n), np.uint64)
    out = dict(sigmas=np.zeros((80, n), np.uint64), wires=np.zeros((135, n), np.uint64), pi_hash=np.zeros(4, np.uint64))
    nc = C.c_uint32(0)
    blob = C.POINTER(C.c_uint64)()
    blen = C.c_size_t(0)
    check(_lib.load().eng_synth_circuit_v2(degree_bits, seed, kinds_mask, ptr(consts), ptr(out["sigmas"]), ptr(out["wires"]),
                                           ptr(out["pi_hash"]), C.byref(nc), C.byref(blob), C.byref(blen)))
    out["constants"] = consts[:nc.value].copy()
    out["blob"] = _take_blob(blob, blen)
    return out


def build_sigmas(degree_bits, num_routed, copies):
    """eng_build_sigmas: sigma polynomial values [num_routed][n] from copy constraints [(row_a, col_a, row_b, col_b), ...]
    (host code: the permutation half of CircuitBuilder::build())."""
    c = np.ascontiguousarray(np.array(copies, dtype=np.uint32).reshape(-1, 4))
    out = np.zeros((num_routed, 1 << degree_bits), np.uint64)
    check(_lib.load().eng_build_sigmas(degree_bits, num_routed, c.ctypes.data_as(C.c_void_p), c.shape[0], ptr(out)))
    return out


def circuit_describe(header12, gates8, digest4):
    """eng_circuit_describe: version-2 description (with bytecode) of a circuit over library gates."""
    h, g_, d = host_u64(header12), host_u64(gates8).reshape(-1, 8), host_u64(digest4)
    blob = C.POINTER(C.c_uint64)()
    blen = C.c_size_t(0)
    check(_lib.load().eng_circuit_describe(ptr(h), ptr(g_), g_.shape[0], ptr(d), C.byref(blob), C.byref(blen)))
    return _take_blob(blob, blen)


def _col_ptrs(cols):
    cols = [host_u64(c) for c in cols]
    return cols, (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])


def verify(circuit_blob, constants_sigmas_cap, public_inputs_hash, proof_blob):
    """CircuitData::verify (host code, no device needed).  Returns None when the proof verifies; raises EngineError
    (ENG_ERR_INVALID, message = the failed check) otherwise, as plonky2's verify() returns Err."""
    b, cap, pi, pr = host_u64(circuit_blob), host_u64(constants_sigmas_cap).ravel(), host_u64(public_inputs_hash), host_u64(proof_blob)
    check(_lib.load().eng_verify(ptr(b), ptr(cap), ptr(pi), ptr(pr), pr.size))


def public_inputs_hash(public_inputs):
    """PoseidonHash::hash_no_pad(public_inputs) -> the 4 words Circuit.prove / verify take (host code)."""
    pis = host_u64(list(public_inputs))
    out = np.zeros(4, np.uint64)
    check(_lib.load().eng_public_inputs_hash(ptr(pis), pis.size, ptr(out)))
    return out


def proof_to_bytes(circuit_blob, proof_blob, public_inputs=()):
    """ProofWithPublicInputs::to_bytes (plonky2's wire format as restated; see include/plonky2_b200.h)."""
    b, pr, pis = host_u64(circuit_blob), host_u64(proof_blob), host_u64(list(public_inputs))
    out = C.POINTER(C.c_uint8)()
    n = C.c_size_t(0)
    check(_lib.load().eng_proof_to_bytes(ptr(b), ptr(pr), pr.size, ptr(pis), pis.size, C.byref(out), C.byref(n)))
    data = bytes(bytearray(out[:n.value]))
    _lib.load().eng_bytes_free(out)
    return data


def proof_from_bytes(circuit_blob, data):
    """ProofWithPublicInputs::from_bytes -> (proof blob, public inputs)."""
    b = host_u64(circuit_blob)
    buf = np.frombuffer(data, np.uint8).copy()
    pb, pi = C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint64)()
    pn, qn = C.c_size_t(0), C.c_size_t(0)
    check(_lib.load().eng_proof_from_bytes(ptr(b), ptr(buf), buf.size, C.byref(pb), C.byref(pn), C.byref(pi), C.byref(qn)))
    proof = _take_blob(pb, pn)
    pis = np.ctypeslib.as_array(pi, shape=(max(qn.value, 1),))[:qn.value].copy()
    _lib.load().eng_blob_free(pi)
    return proof, pis


class Circuit:
    """What the plonk rows need of plonky2's CommonCircuitData / ProverOnlyCircuitData, resident on the device."""

    def __init__(self, blob, constants_sigmas, sigma_values, _handle=None):
        self.blob = host_u64(blob)
        self.constants_sigmas = constants_sigmas
        if _handle is None:
            keep, ptrs = _col_ptrs(sigma_values)
            self._h = C.c_void_p()
            check(_lib.lib().eng_circuit_new(ptr(self.blob), constants_sigmas._o._h, ptrs, C.byref(self._h)))
        else:
            self._h = _handle
        info = _lib.CircuitInfo()
        check(_lib.lib().eng_circuit_info(self._h, C.byref(info)))
        self.info = info
        self.degree_bits, self.num_wires, self.num_routed = info.degree_bits, info.num_wires, info.num_routed_wires
        self.num_selectors, self.num_challenges, self.quotient_degree_factor = info.num_selectors, info.num_challenges, info.quotient_degree_factor
        self.num_gate_constants = info.num_constants - info.num_selectors
        self.rate_bits, self.cap_height, self.num_partial_products = info.rate_bits, info.cap_height, info.num_partial_products

    @classmethod
    def build(cls, synth, rate_bits=3, cap_height=4):
        """The part of CircuitBuilder::build() on the hot path: commit constants || sigmas."""
        cs = PolynomialBatch.from_values(list(synth["constants"]) + list(synth["sigmas"]), rate_bits, False, cap_height)
        return cls(synth["blob"], cs, synth["sigmas"])

    def save(self, path):
        """eng_circuit_save: the prover-data cache (description, constant / sigma polynomials, cap)."""
        check(_lib.lib().eng_circuit_save(self._h, ptr(self.blob), str(path).encode()))

    @classmethod
    def load(cls, path):
        """eng_circuit_load: rebuilds the constants||sigmas commitment on the GPU from the cache file."""
        h = C.c_void_p()
        blob = C.POINTER(C.c_uint64)()
        blen = C.c_size_t(0)
        check(_lib.lib().eng_circuit_load(str(path).encode(), C.byref(h), C.byref(blob), C.byref(blen)))
        b = _take_blob(blob, blen)
        cs = C.c_void_p()
        check(_lib.lib().eng_circuit_constants_sigmas(h, C.byref(cs)))
        owner = _Handle(cs)
        owner.close = lambda: None          # the circuit handle owns this batch
        return cls(b, PolynomialBatch(owner), None, _handle=h)

    def partial_products(self, wire_values, betas, gammas):
        """all_wires_permutation_partial_products -> [num_challenges*(1+num_partial_products)][n], committed order."""
        keep, ptrs = _col_ptrs(wire_values)
        n = 1 << self.degree_bits
        out = np.empty((self.num_challenges * (1 + self.num_partial_products), n), np.uint64)
        b, g_ = host_u64(betas), host_u64(gammas)
        check(_lib.lib().eng_partial_products(self._h, ptrs, ptr(b), ptr(g_), ptr(out)))
        return out

    def quotient(self, wires_batch, zs_pp_batch, public_inputs_hash, betas, gammas, alphas):
        """compute_quotient_polys + split + commit -> PolynomialBatch of num_challenges*quotient_degree_factor chunk polynomials."""
        h = C.c_void_p()
        pi, b, g_, a = host_u64(public_inputs_hash), host_u64(betas), host_u64(gammas), host_u64(alphas)
        check(_lib.lib().eng_quotient(self._h, wires_batch._o._h, zs_pp_batch._o._h, ptr(pi), ptr(b), ptr(g_), ptr(a), C.byref(h)))
        return PolynomialBatch(_Handle(h))

    def prove(self, wire_values, public_inputs_hash):
        """prove_with_partition_witness after witness generation -> (proof blob, stage milliseconds)."""
        keep, ptrs = _col_ptrs(wire_values)
        pi = host_u64(public_inputs_hash)
        blob = C.POINTER(C.c_uint64)()
        n = C.c_size_t(0)
        ms = (C.c_float * 8)()
        check(_lib.lib().eng_prove(self._h, ptrs, ptr(pi), C.byref(blob), C.byref(n), ms))
        out = _take_blob(blob, n)
        names = ("wires commitment", "partial products", "Z commitment", "quotient polys", "quotient commitment", "opening set",
                 "opening proofs (FRI)", "total")
        return out, dict(zip(names, list(ms)))

    def verify(self, public_inputs_hash, proof_blob):
        """data.verify(proof): raises EngineError when the proof is rejected."""
        verify(self.blob, self.constants_sigmas.merkle_tree.cap, public_inputs_hash, proof_blob)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().eng_circuit_free(self._h)
                self._h = None
        except Exception:
            pass
