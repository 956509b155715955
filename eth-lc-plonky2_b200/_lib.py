"""ctypes binding of libplonky2_b200.so (the C ABI of include/plonky2_b200.h).

There is no CPU path: if the shared object is missing it is built with nvcc; if it cannot be loaded, or no
CUDA device is present, the first engine call raises.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

ENG_OK, ENG_ERR_INVALID, ENG_ERR_CUDA, ENG_ERR_OOM, ENG_ERR_STATE = 0, 1, 2, 3, 4


class EngineError(RuntimeError):
    """Non-zero eng_status.  ENG_ERR_INVALID is the analogue of a plonky2 assert!/expect panic."""

    def __init__(self, status, message):
        super().__init__("eng_status %d: %s" % (status, message))
        self.status = status


class BatchInfo(C.Structure):
    _fields_ = [("num_polys", C.c_uint32), ("degree_log", C.c_uint32), ("rate_bits", C.c_uint32),
                ("cap_height", C.c_uint32), ("blinding", C.c_uint32), ("leaf_len", C.c_uint32),
                ("num_leaves", C.c_uint64), ("num_digests", C.c_uint64)]


class CircuitInfo(C.Structure):
    _fields_ = [(k, C.c_uint32) for k in ("degree_bits", "num_wires", "num_routed_wires", "num_constants", "num_selectors", "num_challenges",
                                          "quotient_degree_factor", "num_partial_products", "rate_bits", "cap_height", "num_gates",
                                          "num_gate_constraints")]


_u64p = C.POINTER(C.c_uint64)
_vp = C.c_void_p
_SIGNATURES = {
    "eng_init": [C.c_int32],
    "eng_shutdown": [],
    "eng_last_error": [C.c_char_p, C.c_size_t],
    "eng_set_stream": [_vp],
    "eng_synchronize": [],
    "eng_launch_count": [_u64p],
    "eng_set_option": [C.c_char_p, C.c_int64],
    "eng_reserve": [C.c_size_t],
    "eng_host_register": [_vp, C.c_size_t],
    "eng_host_unregister": [_vp],
    "eng_measure_int_peak": [C.POINTER(C.c_double)],
    "eng_poseidon_permute": [_vp, _vp, C.c_size_t],
    "eng_hash_n": [_vp, C.c_size_t, C.c_size_t, C.c_int32, _vp],
    "eng_two_to_one": [_vp, C.c_size_t, _vp],
    "eng_batch_from_values": [C.POINTER(_vp), C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, C.c_uint64, C.c_uint32, C.POINTER(_vp)],
    "eng_batch_from_coeffs": [C.POINTER(_vp), C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, C.c_uint64, C.c_uint32, C.POINTER(_vp)],
    "eng_batch_from_values_dev": [_vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, C.c_uint64, C.c_uint32, C.POINTER(_vp)],
    "eng_batch_from_coeffs_dev": [_vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, C.c_uint64, C.c_uint32, C.POINTER(_vp)],
    "eng_batch_free": [_vp],
    "eng_release_cached": [],
    "eng_lde_dev": [_vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, C.c_uint32, _vp, _vp],
    "eng_lde_peer_dev": [_vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, C.c_uint32, _vp, _vp, C.POINTER(_vp), C.c_uint32],
    "eng_lde_peer_host": [C.POINTER(_vp), C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, C.c_uint32, _vp, _vp, C.POINTER(_vp), C.c_uint32],
    "eng_peer_buffer_alloc": [C.c_uint64, C.POINTER(_vp), C.c_char_p],
    "eng_peer_buffer_open": [C.c_char_p, C.POINTER(_vp)],
    "eng_peer_buffer_close": [_vp],
    "eng_peer_buffer_free": [_vp],
    "eng_merkle_new": [_vp, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(_vp)],
    "eng_merkle_new_dev": [_vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(_vp)],
    "eng_batch_info": [_vp, C.POINTER(BatchInfo)],
    "eng_batch_cap": [_vp, _vp],
    "eng_batch_digests": [_vp, _vp],
    "eng_batch_coeffs": [_vp, C.c_uint32, _vp],
    "eng_batch_leaves": [_vp, C.c_uint64, C.c_uint64, _vp],
    "eng_batch_lde_values": [_vp, C.c_uint64, C.c_uint64, _vp],
    "eng_batch_merkle_path": [_vp, C.c_uint64, _vp, C.POINTER(C.c_uint32)],
    "eng_batch_device_ptrs": [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)],
    "eng_batch_stage_ms": [_vp, C.POINTER(C.c_float)],
    "eng_challenger_new": [C.POINTER(_vp)],
    "eng_challenger_free": [_vp],
    "eng_challenger_observe": [_vp, _vp, C.c_size_t],
    "eng_challenger_get_challenges": [_vp, _vp, C.c_size_t],
    "eng_challenger_get_state": [_vp, _vp],
    "eng_challenger_set_state": [_vp, _vp],
    "eng_batch_eval_ext": [_vp, _vp, _vp],
    "eng_fri_prove_openings": [_vp, C.POINTER(_vp), C.c_uint32, _vp, C.POINTER(C.c_int32), C.POINTER(_u64p), C.POINTER(C.c_size_t)],
    "eng_blob_free": [_u64p],
    "eng_circuit_new": [_vp, _vp, C.POINTER(_vp), C.POINTER(_vp)],
    "eng_circuit_free": [_vp],
    "eng_build_sigmas": [C.c_uint32, C.c_uint32, _vp, C.c_size_t, _vp],
    "eng_circuit_new_sharded": [_vp, C.POINTER(_vp), C.POINTER(_vp)],
    "eng_partial_products_dev": [_vp, C.POINTER(_vp), _vp, _vp, _vp],
    "eng_partial_products_from_dev": [_vp, _vp, _vp, _vp, _vp],
    "eng_h2d_columns": [C.POINTER(_vp), C.c_uint32, C.c_uint64, _vp],
    "eng_quotient_values_shard_dev": [_vp, _vp, _vp, _vp, C.c_uint32, C.c_uint32, _vp, _vp, _vp, _vp, _vp],
    "eng_quotient_coeffs_from_shards_dev": [_vp, _vp, C.c_uint32, _vp],
    "eng_eval_ext_dev": [_vp, C.c_uint32, C.c_uint32, _vp, _vp],
    "eng_fri_combine_shard_dev": [_vp, C.POINTER(_vp), C.POINTER(C.c_uint32), C.c_uint32, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp],
    "eng_fri_prove_from_layer_dev": [_vp, _vp, _vp, C.POINTER(_u64p), C.POINTER(C.c_size_t), _vp],
    "eng_circuit_info": [_vp, C.POINTER(CircuitInfo)],
    "eng_circuit_describe": [_vp, _vp, C.c_uint32, _vp, C.POINTER(_u64p), C.POINTER(C.c_size_t)],
    "eng_circuit_save": [_vp, _vp, C.c_char_p],
    "eng_circuit_load": [C.c_char_p, C.POINTER(_vp), C.POINTER(_u64p), C.POINTER(C.c_size_t)],
    "eng_circuit_constants_sigmas": [_vp, C.POINTER(_vp)],
    "eng_verify": [_vp, _vp, _vp, _vp, C.c_size_t],
    "eng_proof_to_bytes": [_vp, _vp, C.c_size_t, _vp, C.c_size_t, C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_size_t)],
    "eng_proof_from_bytes": [_vp, _vp, C.c_size_t, C.POINTER(_u64p), C.POINTER(C.c_size_t), C.POINTER(_u64p), C.POINTER(C.c_size_t)],
    "eng_bytes_free": [C.POINTER(C.c_uint8)],
    "eng_public_inputs_hash": [_vp, C.c_size_t, _vp],
    "eng_synth_circuit_v2": [C.c_uint32, C.c_uint64, C.c_uint64, _vp, _vp, _vp, _vp, C.POINTER(C.c_uint32), C.POINTER(_u64p), C.POINTER(C.c_size_t)],
    "eng_partial_products": [_vp, C.POINTER(_vp), _vp, _vp, _vp],
    "eng_quotient": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(_vp)],
    "eng_prove": [_vp, C.POINTER(_vp), _vp, C.POINTER(_u64p), C.POINTER(C.c_size_t), C.POINTER(C.c_float)],
    "eng_synth_circuit": [C.c_uint32, C.c_uint64, _vp, _vp, _vp, _vp, _vp],
}

_lib = None
_initialised_device = None


def so_path():
    return _build.SO


def load(build_if_missing=True):
    """dlopen the engine (building it first if the .so is absent)."""
    global _lib
    if _lib is None:
        if build_if_missing:
            _build.build()
        lib = C.CDLL(_build.SO)
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.argtypes = argtypes
            fn.restype = C.c_int32
        _lib = lib
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def last_error():
    buf = C.create_string_buffer(1024)
    load().eng_last_error(buf, len(buf))
    return buf.value.decode(errors="replace")


def check(status):
    if status != ENG_OK:
        raise EngineError(status, last_error())


def init(device=None):
    """eng_init on `device` (default: LOCAL_RANK, else 0).  Raises EngineError without a CUDA device."""
    global _initialised_device
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if _initialised_device is None:
        check(load().eng_init(device))
        _initialised_device = device
    elif _initialised_device != device:
        raise EngineError(ENG_ERR_STATE, "engine already initialised on device %d" % _initialised_device)
    return _initialised_device


def lib():
    """The loaded library with the engine initialised."""
    init()
    return _lib


def set_stream(cuda_stream_handle):
    check(lib().eng_set_stream(_vp(cuda_stream_handle or 0)))


def set_option(name, value):
    """eng_set_option: A/B switches of the engine (see include/plonky2_b200.h)."""
    check(load().eng_set_option(name.encode(), int(value)))


def reserve(num_bytes):
    """eng_reserve: grow the device memory pool ahead of the first proof."""
    check(lib().eng_reserve(int(num_bytes)))


def host_register(array):
    """eng_host_register: page-lock a numpy array (witness columns) for PCIe-speed column copies."""
    check(lib().eng_host_register(C.c_void_p(array.ctypes.data), array.nbytes))


def host_unregister(array):
    check(lib().eng_host_unregister(C.c_void_p(array.ctypes.data)))


def release_cached():
    """eng_release_cached: return cached device buffers and the pool's unused memory to the driver."""
    check(lib().eng_release_cached())


def synchronize():
    check(lib().eng_synchronize())


def launch_count():
    out = C.c_uint64(0)
    check(lib().eng_launch_count(C.byref(out)))
    return out.value


def measure_int_peak():
    """Measured integer issue rates: {'imad', 'imad_wide', 'alu'} in thread-ops/s."""
    out = (C.c_double * 3)()
    check(lib().eng_measure_int_peak(out))
    return {"imad": out[0], "imad_wide": out[1], "alu": out[2]}


def host_u64(a):
    """C-contiguous uint64 view/copy of array-like `a`."""
    return np.ascontiguousarray(a, dtype=np.uint64)


def ptr(a):
    return a.ctypes.data_as(_vp)
