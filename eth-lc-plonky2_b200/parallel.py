"""Multi-GPU commit: column-sharded iNTT/LDE -> one all-to-all -> row-sharded Poseidon hashing -> cap all-gather.

SURVEY.md 8(e).  The reference is single-process (rayon threads inside plonky2); this is the B200 design for
PolynomialBatch::from_values across the GPUs of one box, one process per GPU:

  1. rank r owns a contiguous block of columns; iNTT and coset LDE are independent per column (eng_lde_dev);
     the LDE is written directly as [G][C_r][L/G] so that slice g -- LDE rows [g*L/G, (g+1)*L/G) in bit-reversed
     order, i.e. the leaves of row-shard owner g -- is one contiguous send chunk;
  2. one all-to-all (NCCL over NVLink/NVSwitch; gloo in the CPU tests) turns column shards into row shards:
     afterwards rank g holds [C][L/G], every column of its L/G leaves;
  3. rank g hashes its leaves and builds its 2^(cap_height - log2 G) cap sub-trees (eng_merkle_new_dev); the
     `digests` of the global tree are the concatenation of the per-rank digests in rank order;
  4. a (2^cap_height x 32 B) all-gather replicates the cap.

Fused exchange (PeerExchange, the default on GPUs): steps 1 and 2 are ONE kernel.  Every rank exports its [C][L/G] leaf
matrix over CUDA IPC; the last pass of the LDE stores row shard g straight into rank g's matrix through the peer
mapping (NVLink P2P stores issued tile by tile as the transform finishes them), so no send buffer is written, re-read
or copied and NCCL only carries the two barriers and the cap all-gather.  The NCCL all-to-all path stays for gloo (CPU
tests) and as the A/B baseline.

Coefficients stay column-sharded, leaves and digests row-sharded.  The local operators are injected (`ops`) so that
the host-side index logic can be exercised on CPU with gloo; the default operators call the CUDA engine.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import EngineError, check


def _log2_strict(x, what):
    if x <= 0 or x & (x - 1):
        raise EngineError(_lib.ENG_ERR_INVALID, "%s = %d is not a power of two" % (what, x))
    return x.bit_length() - 1


class ShardPlan:
    """Who owns what for a batch of `num_polys` columns x 2^log_n rows over `world` ranks."""

    def __init__(self, num_polys, log_n, rate_bits, cap_height, world):
        self.num_polys, self.log_n, self.rate_bits, self.cap_height, self.world = num_polys, log_n, rate_bits, cap_height, world
        self.log_world = _log2_strict(world, "world size")
        self.log_l = log_n + rate_bits
        if self.log_world > cap_height:
            raise EngineError(_lib.ENG_ERR_INVALID, "world size %d needs cap_height >= %d (each rank owns whole cap sub-trees)"
                              % (world, self.log_world))
        if cap_height > self.log_l:
            raise EngineError(_lib.ENG_ERR_INVALID, "cap_height %d > log2(leaves) %d" % (cap_height, self.log_l))
        if num_polys < world:
            raise EngineError(_lib.ENG_ERR_INVALID, "fewer columns (%d) than ranks (%d)" % (num_polys, world))
        base, extra = divmod(num_polys, world)
        self.col_counts = [base + (1 if r < extra else 0) for r in range(world)]
        self.col_offsets = [sum(self.col_counts[:r]) for r in range(world)]
        self.rows_per_rank = (1 << self.log_l) >> self.log_world
        self.local_cap_height = cap_height - self.log_world

    def columns_of(self, rank):
        return range(self.col_offsets[rank], self.col_offsets[rank] + self.col_counts[rank])

    def owner_of_leaf(self, leaf_index):
        return leaf_index // self.rows_per_rank, leaf_index % self.rows_per_rank

    def send_splits(self, rank):
        return [self.col_counts[rank] * self.rows_per_rank] * self.world

    def recv_splits(self):
        return [c * self.rows_per_rank for c in self.col_counts]


class EngineOps:
    """Local operators backed by the CUDA engine (torch tensors are device memory only)."""

    def __init__(self, device):
        import torch
        self.torch, self.device = torch, device

    def empty(self, numel):
        return self.torch.empty(numel, dtype=self.torch.int64, device=self.device)

    def lde(self, src, is_values, log_n, rate_bits, log_row_shards, coeffs_out, lde_out):
        num_polys = src.shape[0]
        check(_lib.lib().eng_lde_dev(C.c_void_p(src.data_ptr()), num_polys, log_n, rate_bits, int(is_values), log_row_shards,
                                     C.c_void_p(coeffs_out.data_ptr()), C.c_void_p(lde_out.data_ptr())))
        _lib.synchronize()   # the exchange runs on torch's stream

    def after_exchange(self):
        """The exchange (all_to_all_single / copy_) was enqueued on torch's current stream and returns at once; the leaf
        hashing runs on the ENGINE's stream, which has no dependency on it.  Wait for the exchange before hashing."""
        self.torch.cuda.current_stream(self.device).synchronize()

    def merkle(self, rows_colmajor, num_polys, num_rows, cap_height):
        from .plonky2 import MerkleTree, _Handle
        h = C.c_void_p()
        check(_lib.lib().eng_merkle_new_dev(C.c_void_p(rows_colmajor.data_ptr()), 1, num_rows, num_rows, num_polys, cap_height, C.byref(h)))
        return MerkleTree(_Handle(h))

    def to_tensor(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).to(self.device)

    def to_numpy(self, t):
        return t.cpu().numpy().view(np.uint64)


class _DevArray:
    """Exposes raw device memory to torch (torch.as_tensor) through __cuda_array_interface__."""

    def __init__(self, ptr, numel):
        self.__cuda_array_interface__ = {"shape": (numel,), "typestr": "<i8", "data": (ptr, False), "version": 2}


class PeerExchange:
    """Receive side of the fused exchange for one ShardPlan: this rank's [C][L/G] leaf matrices (allocated outside the pool,
    exported over CUDA IPC) and the peer mappings of everybody else's.  Collective constructor; reusable across steps.

    `slots` leaf matrices form a ring: a batch built in slot s keeps referring to that matrix (its `rows` and the leaves
    behind MerkleTree::get are views of it, not copies), so a prover that keeps several oracles alive (wires, Z, quotient
    ... until the FRI query phase) gives each its own slot.  Every from_values bumps the slot's generation; accessors of an
    older batch of the same slot raise instead of returning rows that no longer match their digests."""

    def __init__(self, plan, rank, device, group=None, dist=None, slots=1):
        import torch
        if dist is None:
            import torch.distributed as dist
        self.plan, self.rank, self.group, self.dist = plan, rank, group, dist
        lib = _lib.lib()
        if slots < 1:
            raise EngineError(_lib.ENG_ERR_INVALID, "PeerExchange needs at least one slot")
        self.slots = slots
        self.slot_elems = plan.num_polys * plan.rows_per_rank
        self.generation = [0] * slots
        elems = self.slot_elems * slots
        self.local_ptr, self.bases, self._opened = None, [], []
        self._token = torch.zeros(1, dtype=torch.int32, device=device)
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        # every step is followed by an agreement (all-reduce of a failure count): a rank that cannot allocate or cannot
        # map a peer must not leave the others waiting inside a collective
        err = None
        try:
            check(lib.eng_peer_buffer_alloc(elems, C.byref(ptr), handle))
            self.local_ptr = ptr.value
        except EngineError as e:
            err = e
        self._agree(err, "allocating the exchange buffer")
        handles = [None] * plan.world
        dist.all_gather_object(handles, handle.raw, group=group)
        try:
            for g in range(plan.world):
                if g == rank:
                    self.bases.append(self.local_ptr)
                else:
                    q = C.c_void_p()
                    check(lib.eng_peer_buffer_open(handles[g], C.byref(q)))
                    self.bases.append(q.value)
                    self._opened.append(q.value)
        except EngineError as e:
            err = e
        self._agree(err, "mapping the peers' exchange buffers (CUDA IPC / peer access)")
        off = plan.col_offsets[rank] * plan.rows_per_rank * 8     # this rank's first column inside every leaf matrix
        self._shard_out = [(C.c_void_p * plan.world)(*[b + s * self.slot_elems * 8 + off for b in self.bases]) for s in range(slots)]
        self._recv = torch.as_tensor(_DevArray(self.local_ptr, elems), device=device).view(slots, self.slot_elems)

    @property
    def shard_out(self):
        return self._shard_out[0]

    @property
    def recv(self):
        return self._recv[0]

    def slot_shard_out(self, slot):
        return self._shard_out[slot]

    def slot_recv(self, slot):
        return self._recv[slot]

    def _agree(self, err, what):
        """Collective: raises on EVERY rank if any rank failed."""
        self._token.fill_(1 if err is not None else 0)
        self.dist.all_reduce(self._token, group=self.group)
        failed = int(self._token.item())
        self._token.zero_()
        if failed:
            self._release_local()
            raise EngineError(_lib.ENG_ERR_CUDA, "fused exchange unavailable: %d rank(s) failed %s%s" % (failed, what, ": %s" % err if err else ""))

    def _release_local(self):
        lib = _lib.lib()
        for q in self._opened:
            lib.eng_peer_buffer_close(C.c_void_p(q))
        self._opened = []
        if self.local_ptr:
            lib.eng_peer_buffer_free(C.c_void_p(self.local_ptr))
            self.local_ptr = None

    def barrier(self):
        self.dist.all_reduce(self._token, group=self.group)
        self._token.zero_()
        import torch
        torch.cuda.synchronize()

    def close(self):
        lib = _lib.lib()
        self.barrier()
        for q in self._opened:
            lib.eng_peer_buffer_close(C.c_void_p(q))
        self._opened = []
        self.barrier()
        if self.local_ptr:
            lib.eng_peer_buffer_free(C.c_void_p(self.local_ptr))
            self.local_ptr = None


class ShardedPolynomialBatch:
    """PolynomialBatch whose polynomials are column-sharded and whose leaves / digests are row-sharded."""

    def __init__(self, plan, rank, coeffs, rows, tree, cap, ops, exchange=None, slot=0):
        self.plan, self.rank, self.coeffs, self._rows, self.merkle_tree_local, self.cap, self.ops = plan, rank, coeffs, rows, tree, cap, ops
        # with a PeerExchange the leaves live in the exchange's slot (not owned): valid until the slot is reused
        self._exchange, self._slot = exchange, slot
        self._generation = exchange.generation[slot] if exchange is not None else 0

    def _check_live(self):
        if self._exchange is not None and self._exchange.generation[self._slot] != self._generation:
            raise EngineError(_lib.ENG_ERR_STATE, "the leaf matrix of this batch (exchange slot %d) was overwritten by a later "
                              "from_values; give every live batch its own slot (PeerExchange(slots=k))" % self._slot)

    @property
    def rows(self):
        """This rank's leaves, column-major [C][L/G] (a view of the exchange slot when the fused exchange built them)."""
        self._check_live()
        return self._rows

    @classmethod
    def from_values(cls, local_values, plan, rank, group=None, ops=None, is_values=True, dist=None, exchange=None, slot=0):
        """local_values: [plan.col_counts[rank]][2^log_n] tensor holding this rank's columns (values, or coefficients
        when is_values is False).  Collective over `group` (torch.distributed).  With `exchange` (a PeerExchange of the
        same plan) the column->row exchange is fused into the LDE's last pass; the leaves land in leaf matrix `slot` of the
        exchange, which the returned batch refers to until a later call reuses that slot."""
        if dist is None:
            import torch.distributed as dist
        if ops is None and not isinstance(local_values, (list, tuple)):
            ops = EngineOps(local_values.device)
        n = 1 << plan.log_n
        c_r = plan.col_counts[rank]
        host_cols = None
        if isinstance(local_values, (list, tuple)):       # host columns (numpy uint64): only with the fused exchange
            if exchange is None or plan.world == 1:
                raise EngineError(_lib.ENG_ERR_INVALID, "host columns need exchange= (PeerExchange) and world > 1")
            host_cols = [np.ascontiguousarray(c, dtype=np.uint64) for c in local_values]
            if len(host_cols) != c_r or any(c.shape != (n,) for c in host_cols):
                raise EngineError(_lib.ENG_ERR_INVALID, "rank %d expects %d host columns of %d elements" % (rank, c_r, n))
            if ops is None:
                ops = EngineOps(exchange.recv.device)
        elif tuple(local_values.shape) != (c_r, n):
            raise EngineError(_lib.ENG_ERR_INVALID, "rank %d expects a [%d][%d] column shard, got %s" % (rank, c_r, n, tuple(local_values.shape)))
        coeffs = ops.empty(c_r * n).view(c_r, n)
        if exchange is not None and plan.world > 1:
            import os, time
            trace = os.environ.get("ENG_TRACE")
            t0 = time.perf_counter()
            if not 0 <= slot < exchange.slots:
                raise EngineError(_lib.ENG_ERR_INVALID, "exchange slot %d out of range (%d slots)" % (slot, exchange.slots))
            scratch = ops.empty(c_r * (n << plan.rate_bits))
            exchange.generation[slot] += 1           # earlier batches of this slot are dead from here on
            shard_out = exchange.slot_shard_out(slot)
            exchange.barrier()                       # every rank has finished reading its leaf matrix of the previous call
            t1 = time.perf_counter()
            if host_cols is not None:
                ptrs = (C.c_void_p * c_r)(*[c.ctypes.data for c in host_cols])
                check(_lib.lib().eng_lde_peer_host(ptrs, c_r, plan.log_n, plan.rate_bits, int(is_values), plan.log_world,
                                                   C.c_void_p(coeffs.data_ptr()), C.c_void_p(scratch.data_ptr()), shard_out, rank))
            else:
                check(_lib.lib().eng_lde_peer_dev(C.c_void_p(local_values.data_ptr()), c_r, plan.log_n, plan.rate_bits, int(is_values),
                                                  plan.log_world, C.c_void_p(coeffs.data_ptr()), C.c_void_p(scratch.data_ptr()),
                                                  shard_out, rank))
            _lib.synchronize()
            t2 = time.perf_counter()
            exchange.barrier()                       # every rank's stores have landed
            t3 = time.perf_counter()
            del scratch
            recv = exchange.slot_recv(slot)
            tree = ops.merkle(recv, plan.num_polys, plan.rows_per_rank, plan.local_cap_height)
            local_cap = ops.to_tensor(tree.cap).reshape(-1)
            t4 = time.perf_counter()
            parts = [ops.empty(local_cap.numel()) for _ in range(plan.world)]
            dist.all_gather(parts, local_cap, group=group)
            cap = np.concatenate([ops.to_numpy(p).reshape(-1, 4) for p in parts])
            if trace:
                print("rank %d: alloc+barrier %.1f  lde %.1f  barrier %.1f  merkle+cap %.1f  gather %.1f ms" % (
                    rank, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (time.perf_counter() - t4)), flush=True)
            return cls(plan, rank, coeffs, recv.view(plan.num_polys, plan.rows_per_rank), tree, cap, ops, exchange, slot)
        send = ops.empty(c_r * (n << plan.rate_bits))
        ops.lde(local_values, is_values, plan.log_n, plan.rate_bits, plan.log_world, coeffs, send)   # [G][C_r][L/G]
        recv = ops.empty(plan.num_polys * plan.rows_per_rank)                                          # [C][L/G]
        if plan.world > 1:
            dist.all_to_all_single(recv, send, output_split_sizes=plan.recv_splits(), input_split_sizes=plan.send_splits(rank), group=group)
        else:
            recv.copy_(send)
        getattr(ops, "after_exchange", lambda: None)()     # torch stream -> engine stream ordering (see EngineOps)
        del send
        tree = ops.merkle(recv, plan.num_polys, plan.rows_per_rank, plan.local_cap_height)
        local_cap = ops.to_tensor(tree.cap).reshape(-1)
        if plan.world > 1:
            parts = [ops.empty(local_cap.numel()) for _ in range(plan.world)]
            dist.all_gather(parts, local_cap, group=group)
            cap = np.concatenate([ops.to_numpy(p).reshape(-1, 4) for p in parts])
        else:
            cap = ops.to_numpy(local_cap).reshape(-1, 4)
        return cls(plan, rank, coeffs, recv.view(plan.num_polys, plan.rows_per_rank), tree, cap, ops)

    # ---- accessors for data this rank owns ----
    def owns_leaf(self, leaf_index):
        return self.plan.owner_of_leaf(leaf_index)[0] == self.rank

    def get(self, leaf_index):
        """merkle_tree.get(leaf_index) (only on the owning rank)."""
        owner, local = self.plan.owner_of_leaf(leaf_index)
        if owner != self.rank:
            raise EngineError(_lib.ENG_ERR_INVALID, "leaf %d lives on rank %d" % (leaf_index, owner))
        self._check_live()
        return self.merkle_tree_local.get(local)

    def prove(self, leaf_index):
        """merkle_tree.prove(leaf_index).siblings (only on the owning rank); same siblings as the global tree."""
        owner, local = self.plan.owner_of_leaf(leaf_index)
        if owner != self.rank:
            raise EngineError(_lib.ENG_ERR_INVALID, "leaf %d lives on rank %d" % (leaf_index, owner))
        return self.merkle_tree_local.prove(local)

    @property
    def local_digests(self):
        """This rank's slice of the global `digests` vector (slices concatenate in rank order)."""
        return self.merkle_tree_local.digests
