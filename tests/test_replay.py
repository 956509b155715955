"""CPU replay of the CUDA kernel bodies (tests/emu) against the oracle.

The dev container has no GPU; tests/emu compiles the __host__ __device__ bodies of the kernels with g++ and replays
each launch block by block / thread by thread.  This checks tile index maps, twiddle and shift tables, the pass
planner and the Merkle digest layout before the `-m gpu` tests run the real kernels through the C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import P, rand_field

HERE = os.path.dirname(os.path.abspath(__file__))
u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def emu():
    # EMU_ASAN=1 (tools/asan_replay.sh): the replay harness built with AddressSanitizer -- every load / store of the kernel
    # bodies (tile staging, four-step index maps, leaf matrices of the row shards, digests) is bounds-checked on the CPU.
    asan = os.environ.get("EMU_ASAN") == "1"
    so = os.path.join(HERE, "emu", "libemu_asan.so" if asan else "libemu.so")
    src = os.path.join(HERE, "emu", "emu.cpp")
    csrc = os.path.join(HERE, "..", "eth-lc-plonky2_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        flags = ["-O1", "-g", "-fsanitize=address", "-fno-omit-frame-pointer"] if asan else ["-O2"]
        subprocess.check_call(["/usr/bin/g++"] + flags + ["-std=c++17", "-fPIC", "-shared", "-o", so, src])
    L = C.CDLL(so)
    L.emu_poseidon_permute.argtypes = [u64p, u64p, C.c_size_t]
    L.emu_poseidon_permute_f64.argtypes = [u64p, u64p, C.c_size_t]
    L.emu_poseidon_f64_ranges.argtypes = [C.POINTER(C.c_double), C.c_int]
    L.emu_batch_from_values.argtypes = [u64p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_uint64, u64p, u64p, u64p, u64p]
    L.emu_batch_from_values.restype = C.c_int
    L.emu_plan.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, u64p, C.c_int]
    L.emu_plan.restype = C.c_int
    for f in (L.emu_gl_mul, L.emu_gl_add, L.emu_gl_sub):
        f.argtypes = [C.c_uint64, C.c_uint64]; f.restype = C.c_uint64
    return L


def test_field_ops_any_u64(emu):
    rng = np.random.default_rng(5)
    edge = [0, 1, P - 1, P, P + 1, 2**64 - 1, 2**32 - 1, 2**32, 2**64 - 2**32]
    pairs = [(a, b) for a in edge for b in edge] + [tuple(int(v) for v in x) for x in rand_field(rng, (300, 2), noncanonical=True)]
    for a, b in pairs:
        assert emu.emu_gl_mul(a, b) == (a * b) % P
        assert emu.emu_gl_add(a, b) == (a + b) % P
        assert emu.emu_gl_sub(a, b) == (a - b) % P


def test_poseidon_body(emu, oracle):
    rng = np.random.default_rng(1)
    st = rand_field(rng, (40, 12), noncanonical=True)
    st[0] = 0; st[1] = np.arange(12); st[2] = P - 1; st[3] = 2**64 - 1
    out = np.zeros_like(st)
    emu.emu_poseidon_permute(st, out, st.shape[0])
    for i in range(st.shape[0]):
        assert (oracle.poseidon(st[i]) == out[i]).all()


def test_poseidon_fp64_formulation(emu, oracle):
    """poseidon_f64.cuh (the device permutation: linear layers in exact binary64, partial rounds two at a time) replayed
    with host IEEE doubles: the four upstream KATs, edge inputs, and agreement with the integer formulation."""
    from helpers import golden, unhx
    rng = np.random.default_rng(7)
    st = rand_field(rng, (3000, 12), noncanonical=True)
    st[0] = 0; st[1] = np.arange(12); st[2] = P - 1; st[3] = 2**64 - 1; st[4] = 2**32 - 1; st[5] = 2**64 - 2**32
    st[6] = np.array([0, 2**64 - 1] * 6, np.uint64); st[7] = P
    out = np.zeros_like(st); ref = np.zeros_like(st)
    emu.emu_poseidon_permute_f64(st, out, st.shape[0])
    emu.emu_poseidon_permute(st, ref, st.shape[0])
    assert (out == ref).all()
    for i in range(8):
        assert (oracle.poseidon(st[i]) == out[i]).all()
    assert int(out[0][0]) == 0x3C18A9786CB0B359      # upstream KAT, all-zero input


CASES = [(3, 0, 1, 0), (3, 1, 1, 1), (5, 2, 2, 0), (9, 3, 1, 1), (3, 3, 3, 2), (135, 4, 3, 4), (7, 5, 3, 2), (9, 7, 2, 3),
         (2, 10, 1, 4), (3, 12, 2, 4), (3, 13, 1, 4), (2, 13, 3, 0), (3, 14, 2, 4), (1, 15, 1, 2)]   # 15: odd two-pass split (7 + 8)


@pytest.mark.parametrize("C_,log_n,r,h", CASES)
@pytest.mark.parametrize("is_values", [1, 0])
def test_commit_replay_matches_oracle(emu, oracle, C_, log_n, r, h, is_values):
    if not is_values and log_n not in (3, 7, 13):
        pytest.skip("from_coeffs replay on a subset")
    rng = np.random.default_rng(1000 * C_ + log_n)
    n = 1 << log_n; L = n << r
    vals = rand_field(rng, (C_, n), noncanonical=True)
    coeffs = np.zeros((C_, n), np.uint64); lde = np.zeros((C_, L), np.uint64)
    nd = 2 * (L - (1 << h))
    dig = np.zeros((max(nd, 1), 4), np.uint64); cap = np.zeros((1 << h, 4), np.uint64)
    assert emu.emu_batch_from_values(vals, C_, log_n, r, h, is_values, 3, coeffs, lde, dig, cap) == 0
    b = oracle.Batch.from_values(vals, r, h) if is_values else oracle.Batch.from_coeffs(vals, r, h)
    assert (b.coeffs == coeffs).all()
    assert (b.leaves == lde.T).all()          # engine layout is the transpose: column-major, bit-reversed rows
    assert nd == 0 or (b.digests == dig[:nd]).all()
    assert (b.cap == cap).all()


@pytest.mark.parametrize("log_n", range(0, 27))
def test_planner_shapes(emu, log_n):
    out = np.zeros(6 * 4, np.uint64)
    for intt in (0, 1):
        k = emu.emu_plan(135, log_n, 3, intt, out, 4)
        assert k == (1 if log_n <= 13 else 2)
        for mode, log_p, log_a, threads, smem, tiles in out.reshape(4, 6)[:k]:
            assert log_p <= 13 and threads in range(32, 513) and smem <= 227 * 1024 and tiles > 0
            if mode in (0, 3, 4):                  # strided passes: at least 32-byte segments (16 for 2^13-point tiles)
                assert log_a >= (1 if log_p == 13 else 2)


@pytest.mark.parametrize("C_,log_n,r,log_g", [(5, 4, 3, 1), (5, 4, 3, 3), (3, 13, 3, 3), (3, 13, 3, 2), (2, 13, 1, 3), (7, 6, 2, 3)])
def test_row_sharded_lde_layout(emu, oracle, C_, log_n, r, log_g):
    """[G][C][L/G]: slice g is the contiguous all-to-all chunk holding LDE rows [g*L/G, (g+1)*L/G) of every column."""
    emu.emu_batch_sharded.argtypes = [u64p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_uint64, C.c_uint32, u64p, u64p, u64p, u64p]
    emu.emu_batch_sharded.restype = C.c_int
    rng = np.random.default_rng(C_ + log_n)
    n = 1 << log_n; L = n << r; G = 1 << log_g
    vals = rand_field(rng, (C_, n))
    coeffs = np.zeros((C_, n), np.uint64); lde = np.zeros((G, C_, L // G), np.uint64)
    dummy = np.zeros((1, 4), np.uint64)
    assert emu.emu_batch_sharded(vals, C_, log_n, r, 0, 1, 3, log_g, coeffs, lde, dummy, dummy) == 0
    b = oracle.Batch.from_values(vals, r, 0)
    for g in range(G):
        assert (lde[g].T == b.leaves[g * (L // G):(g + 1) * (L // G)]).all()


@pytest.mark.parametrize("C_,log_n,r,log_g", [(5, 4, 3, 1), (7, 6, 2, 3), (5, 13, 3, 2), (6, 14, 1, 3), (19, 9, 3, 4)])
def test_fused_exchange_addressing(emu, oracle, C_, log_n, r, log_g):
    """CPU replay of eng_lde_peer_dev: every emulated rank transforms its column block and the LAST pass of the LDE stores
    row shard g through shard_out[g] (here: plain host matrices standing in for the peer mappings), walking the shards
    from its own rank on.  Afterwards matrix g must hold rows [g L/G, (g+1) L/G) of the oracle's leaves, all columns."""
    G = 1 << log_g
    n = 1 << log_n; L = n << r; rows = L // G
    rng = np.random.default_rng(31 * C_ + log_n)
    vals = rand_field(rng, (C_, n), noncanonical=True)
    ref = oracle.Batch.from_values(vals, r, 0)
    mats = [np.zeros((C_, rows), np.uint64) for _ in range(G)]
    ptrs = (C.c_void_p * G)(*[m.ctypes.data for m in mats])
    emu.emu_lde_peer.argtypes = [u64p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                 C.POINTER(C.c_void_p), u64p, u64p]
    emu.emu_lde_peer.restype = C.c_int
    base, extra = divmod(C_, G)
    col0 = 0
    for rank in range(G):
        c_r = base + (1 if rank < extra else 0)
        if c_r == 0:
            continue
        local = np.ascontiguousarray(vals[col0:col0 + c_r])
        coeffs = np.zeros((c_r, n), np.uint64); scratch = np.zeros((c_r, L), np.uint64)
        assert emu.emu_lde_peer(local, c_r, log_n, r, 3, log_g, rank, col0, ptrs, coeffs, scratch) == 0
        assert (coeffs == ref.coeffs[col0:col0 + c_r]).all()
        col0 += c_r
    for g in range(G):
        assert (mats[g].T == ref.leaves[g * rows:(g + 1) * rows]).all()


def test_poseidon_fp64_magnitudes_stay_exact(emu):
    """The FP64 formulation is exact as long as every value entering a fold lies in [-2^51, 2^51) (the fold reads the
    mantissa of x + 1.5 * 2^52) and the accumulated sums stay far below 2^53.  The CPU replay records those ranges over
    random and adversarial states (all-ones, p - 1, alternating extremes)."""
    rng = np.random.default_rng(11)
    st = rand_field(rng, (20000, 12), noncanonical=True)
    st[0] = 2**64 - 1; st[1] = P - 1; st[2] = np.array([0, 2**64 - 1] * 6, np.uint64); st[3] = np.array([2**64 - 1, 0] * 6, np.uint64)
    st[4] = 2**63; st[5] = 2**32 - 1; st[6] = 2**64 - 2**32; st[7] = 0
    out = np.zeros_like(st)
    r = (C.c_double * 3)()
    emu.emu_poseidon_f64_ranges(r, 1)
    emu.emu_poseidon_permute_f64(st, out, st.shape[0])
    emu.emu_poseidon_f64_ranges(r, 1)
    fold_min, fold_max, renorm_max = r[0], r[1], r[2]
    assert -2.0**51 <= fold_min and fold_max < 2.0**51, (fold_min, fold_max)
    assert fold_max < 2.0**49                      # accumulated sums (observed ~2^48); -2^51 only as the explicit bias
    assert 2.0**40 < renorm_max < 2.0**50, renorm_max


def test_ntt_lazy_values_stay_in_range(emu, oracle):
    """The radix-16 DFT rounds add and subtract LAZILY on 96-bit two's-complement values (no wrap correction) and shift them
    by up to 31 bits into a 128-bit intermediate: that is exact while |v| < 2^80.  The CPU replay records the largest lazy
    magnitude over a commit of all-ones columns (every butterfly input at its maximum) and a random one."""
    emu.emu_l3_maxabs.argtypes = [C.c_int]
    emu.emu_l3_maxabs.restype = C.c_double
    C_, log_n, r = 3, 13, 3
    n = 1 << log_n; L = n << r
    worst = np.full((C_, n), 2**64 - 1, np.uint64)
    worst[1] = P - 1
    worst[2] = rand_field(np.random.default_rng(2), (n,), noncanonical=True)
    coeffs = np.zeros((C_, n), np.uint64); lde = np.zeros((C_, L), np.uint64)
    dig = np.zeros((2 * (L - 1), 4), np.uint64); cap = np.zeros((1, 4), np.uint64)
    emu.emu_l3_maxabs(1)
    assert emu.emu_batch_from_values(worst, C_, log_n, r, 0, 1, 3, coeffs, lde, dig, cap) == 0
    m = emu.emu_l3_maxabs(1)
    assert 2.0**64 < m < 2.0**72, m            # 16 inputs below 2^64 summed: < 2^68; shift results stay below 2^66
    b = oracle.Batch.from_values(worst, r, 0)
    assert (b.coeffs == coeffs).all() and (b.leaves == lde.T).all()
