"""B200-native plonky2 commitment engine for the eth-lc-plonky2 light-client circuit.

The product is libplonky2_b200.so (CUDA kernels for sm_100a behind the C ABI of include/plonky2_b200.h);
this package is the thin host-side mirror of the plonky2 operator surface used by tests, bench.py and the
torch.distributed plumbing.  There is no CPU path.
"""
from . import build as _build_mod
from ._lib import (ENG_ERR_CUDA, ENG_ERR_INVALID, ENG_ERR_OOM, ENG_ERR_STATE, ENG_OK, EngineError, exported_symbols, host_register,
                   host_unregister, init,
                   launch_count, load, measure_int_peak, release_cached, reserve, set_option, set_stream, so_path, synchronize)
from .synthetic import splitmix_columns
from .parallel import EngineOps, PeerExchange, ShardedPolynomialBatch, ShardedProver, ShardPlan, splice_initial_openings
from .plonky2 import ALL_GATES, GATE_KINDS, Circuit, build_sigmas, circuit_describe, proof_from_bytes, proof_to_bytes, public_inputs_hash, synth_circuit, synth_circuit_v2, verify
from .plonky2 import Challenger, FriInstanceInfo, FriParams, FriProof, reduction_arity_bits
from .plonky2 import SALT_SIZE, MerkleTree, PolynomialBatch, PoseidonHash, poseidon


def build(force=False, verbose=False):
    """Compile the CUDA extension in-tree for sm_100a."""
    return _build_mod.build(force=force, verbose=verbose)
