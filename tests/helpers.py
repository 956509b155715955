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


class ThreadDist:
    """Stand-in for torch.distributed with `world` ranks running as THREADS of one process on one GPU: enough of the API
    (all_gather, all_gather_object, all_to_all_single, barrier) for ShardedPolynomialBatch / ShardedProver, so that the
    single-GPU `-m gpu` run exercises the multi-rank host logic and the row-sharded kernels with real collectives'
    semantics.  (The real NCCL / CUDA-IPC path is checked by bench.py --gpus N, see parity_check in its output.)"""

    def __init__(self, world):
        import threading
        self.world = world
        self._bar = threading.Barrier(world, timeout=600)
        self._slots = [None] * world
        self._tls = threading.local()

    def set_rank(self, rank):
        self._tls.rank = rank

    def abort(self):
        self._bar.abort()

    def _exchange(self, item):
        import torch
        self._slots[self._tls.rank] = item
        torch.cuda.synchronize()
        self._bar.wait()
        items = list(self._slots)
        self._bar.wait()
        return items

    def barrier(self, group=None):
        self._bar.wait()

    def all_gather(self, out_list, t, group=None):
        import torch
        items = self._exchange(t)
        for g in range(self.world):
            out_list[g].copy_(items[g])
        torch.cuda.synchronize()
        self._bar.wait()            # nobody overwrites its send tensor before everybody has copied it

    def all_gather_object(self, out_list, obj, group=None):
        items = self._exchange(obj)
        for g in range(self.world):
            out_list[g] = items[g]

    def all_to_all_single(self, recv, send, output_split_sizes=None, input_split_sizes=None, group=None):
        import torch
        r = self._tls.rank
        items = self._exchange((send, list(input_split_sizes)))
        off = 0
        for g in range(self.world):
            src, splits = items[g]
            start, n = sum(splits[:r]), splits[r]
            assert n == output_split_sizes[g]
            recv[off:off + n].copy_(src[start:start + n])
            off += n
        torch.cuda.synchronize()
        self._bar.wait()


def run_ranks(world, fn):
    """Runs fn(rank, dist) on `world` threads; returns the list of results; re-raises the first failure."""
    import threading
    dist = ThreadDist(world)
    out, errs = [None] * world, []

    def work(rank):
        try:
            dist.set_rank(rank)
            out[rank] = fn(rank, dist)
        except BaseException as e:   # noqa: BLE001
            errs.append(e)
            dist.abort()

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if errs:
        real = [e for e in errs if e.__class__.__name__ != "BrokenBarrierError"]
        raise (real or errs)[0]
    return out
