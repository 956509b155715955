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
        import time as _time
        t_enter = _time.perf_counter()
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
                print("rank %d: enter %.1f  alloc+barrier %.1f  lde %.1f  barrier %.1f  merkle+cap %.1f  gather %.1f ms" % (
                    rank, 1e3 * (t0 - t_enter), 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (time.perf_counter() - t4)), flush=True)
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


# ---------------------------------------------------------------------------------------------------------------------
# Multi-GPU prover (include/plonky2_b200.h "multi-GPU prover")
# ---------------------------------------------------------------------------------------------------------------------
P = 0xFFFFFFFF00000001
_POWER_OF_TWO_GENERATOR = 1753635133440165772
_BLOB_V2_MAGIC = 0x32424B4C50


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def splice_initial_openings(fri_blob, openings):
    """fri_blob: FriProof blob of eng_fri_prove_from_layer_dev (0 initial oracles per query round); openings[q] = [(leaf row,
    siblings [layers][4]) per oracle] -> the blob eng_fri_prove_openings would have produced."""
    b = [int(x) for x in fri_blob]
    i = 0
    r = b[i]; i += 1
    for _ in range(r):
        i += 1 + b[i]
    f = b[i]; i += 1 + 2 * f
    i += 1                                   # pow_witness
    q = b[i]; i += 1
    out = b[:i]
    if q != len(openings):
        raise EngineError(_lib.ENG_ERR_INVALID, "%d query rounds, openings for %d" % (q, len(openings)))
    for k in range(q):
        if b[i] != 0:
            raise EngineError(_lib.ENG_ERR_INVALID, "query round %d already carries initial-tree openings" % k)
        i += 1
        out.append(len(openings[k]))
        for leaf, path in openings[k]:
            leaf, path = _u64(leaf).ravel(), _u64(path).reshape(-1, 4)
            out += [leaf.size] + [int(x) for x in leaf] + [path.shape[0]] + [int(x) for x in path.ravel()]
        s = b[i]; j = i + 1
        for _ in range(s):
            j += 1 + 2 * b[j]
            j += 1 + 4 * b[j]
        out += b[i:j]
        i = j
    if i != len(b):
        raise EngineError(_lib.ENG_ERR_INVALID, "trailing words in the FRI blob")
    return np.array(out, np.uint64)


class ShardedProver:
    """prove_with_partition_witness (after witness generation) over the `world` GPUs of one box, one process per GPU.

    Every batch is committed column-sharded / row-sharded (ShardedPolynomialBatch); the quotient's constraint evaluation and
    the FRI combination run on this rank's leaf matrices (x -> w_n x and the 16-point FRI cosets stay inside a row shard);
    the cheap steps are replicated (transcript, partial products, the quotient's coset iNTT, the FRI commit phase on the
    gathered layer 0).  Collectives: the fused column -> row exchange of four commitments, the cap all-gathers, one
    all-gather of the quotient values (num_challenges * L * 8 B in total), one of FRI layer 0 (L * 16 B), two small object
    gathers (openings at zeta, query openings).  The proof is bit-identical to eng_prove's on one GPU."""

    STAGES = ("wires commitment", "partial products", "Z commitment", "quotient polys", "quotient commitment", "opening set",
              "opening proofs (FRI)", "total")

    def __init__(self, blob, constants, sigmas, rank, world, device=None, group=None, dist=None, use_peer=True, ops=None):
        import torch
        self.torch = torch
        if dist is None:
            import torch.distributed as dist
        self.dist, self.group, self.rank, self.world = dist, group, rank, world
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.ops = ops if ops is not None else EngineOps(self.device)
        self.blob = _u64(blob)
        lib = _lib.lib()
        sig = [_u64(c) for c in sigmas]
        ptrs = (C.c_void_p * len(sig))(*[c.ctypes.data for c in sig])
        self._h = C.c_void_p()
        check(lib.eng_circuit_new_sharded(_ptr(self.blob), ptrs, C.byref(self._h)))
        info = _lib.CircuitInfo()
        check(lib.eng_circuit_info(self._h, C.byref(info)))
        self.info = info
        b = [int(x) for x in self.blob]
        v2 = b[0] == _BLOB_V2_MAGIC
        h = b[2:] if v2 else b
        self.pow_bits, self.num_query_rounds, ng = h[9], h[10], h[11]
        self.digest = np.array(b[16:20] if v2 else b[12 + 4 * ng:16 + 4 * ng], np.uint64)
        self.log_world = _log2_strict(world, "world size")
        qdb = (info.quotient_degree_factor - 1).bit_length()
        if qdb != info.rate_bits or self.log_world > qdb:
            raise EngineError(_lib.ENG_ERR_INVALID, "the sharded prover needs quotient_degree_bits == rate_bits and at most 2^%d ranks" % qdb)
        self.n = 1 << info.degree_bits
        self.nch, self.nzs, self.nq = info.num_challenges, info.num_challenges * (1 + info.num_partial_products), info.num_challenges * info.quotient_degree_factor
        self.widths = [info.num_constants + info.num_routed_wires, info.num_wires, self.nzs, self.nq]
        self.plans = [ShardPlan(w, info.degree_bits, info.rate_bits, info.cap_height, world) for w in self.widths]
        self.count = self.plans[0].rows_per_rank
        self.exchanges = [None] * 4
        if use_peer and world > 1:
            self.exchanges = [PeerExchange(p, rank, self.device, group=group, dist=dist) for p in self.plans]
        if len(constants) != info.num_constants or len(sig) != info.num_routed_wires:
            raise EngineError(_lib.ENG_ERR_INVALID, "expected %d constant and %d sigma columns" % (info.num_constants, info.num_routed_wires))
        cols = [_u64(c) for c in constants] + sig
        self.cs = self._commit(0, [cols[c] for c in self.plans[0].columns_of(rank)], True)

    # ---- helpers ----
    def _commit(self, k, local, is_values):
        """local: this rank's columns of batch k -- a list of host columns or a [c_r][n] device tensor."""
        plan, ex = self.plans[k], self.exchanges[k]
        if isinstance(local, (list, tuple)) and ex is None:
            local = self.ops.to_tensor(np.stack(local))
        self._torch_sync()      # a tensor produced on torch's stream is consumed on the engine's stream
        return ShardedPolynomialBatch.from_values(local, plan, self.rank, group=self.group, ops=self.ops, is_values=is_values,
                                                  dist=self.dist, exchange=ex)

    def _torch_sync(self):
        self.torch.cuda.current_stream(self.device).synchronize()

    def _all_gather(self, t):
        if self.world == 1:
            return t
        parts = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(parts, t, group=self.group)
        out = self.torch.cat(parts)
        self._torch_sync()      # the engine reads the gathered tensor on its own stream
        return out

    def _all_gather_object(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.group)
        return out

    def _eval(self, batch, z):
        c_r = batch.coeffs.shape[0]
        out = np.empty((c_r, 2), np.uint64)
        zz = _u64([int(z[0]), int(z[1])])
        check(_lib.lib().eng_eval_ext_dev(C.c_void_p(batch.coeffs.data_ptr()), c_r, self.info.degree_bits, _ptr(zz), _ptr(out)))
        return out

    def fri_instance(self, zeta, gzeta):
        """get_fri_instance: batch 0 = every polynomial of the four oracles at zeta, batch 1 = the Z's at g * zeta."""
        from .plonky2 import FriInstanceInfo
        all_polys = [(o, p) for o, w in enumerate(self.widths) for p in range(w)]
        return FriInstanceInfo([(zeta, all_polys), (gzeta, [(2, p) for p in range(self.nch)])])

    # ---- the proof ----
    def prove(self, wire_cols_host, public_inputs_hash):
        """wire_cols_host: all num_wires witness columns (host; every rank holds them, as every rank ran the generators or
        received the witness).  Returns (proof blob, stage milliseconds); the blob is identical on every rank."""
        import time
        from .plonky2 import Challenger, FriParams
        lib, info, torch = _lib.lib(), self.info, self.torch
        if len(wire_cols_host) != info.num_wires:
            raise EngineError(_lib.ENG_ERR_INVALID, "expected %d wire columns" % info.num_wires)
        wires_host = [_u64(c) for c in wire_cols_host]
        pi = _u64(public_inputs_hash)
        marks = [time.perf_counter()]

        def mark():
            torch.cuda.synchronize()
            marks.append(time.perf_counter())

        ch = Challenger()
        ch.observe_hash(self.digest)
        ch.observe_hash(pi)
        wires = self._commit(1, [wires_host[c] for c in self.plans[1].columns_of(self.rank)], True)
        ch.observe_cap(wires.cap)
        betas, gammas = _u64(ch.get_n_challenges(self.nch)), _u64(ch.get_n_challenges(self.nch))
        mark()
        # partial products: replicated (one pass over the routed wires + a scan over the rows; 6 ms at 2^22 rows).  Every rank
        # uploads 1/G of the routed wire columns and the ranks all-gather them over NVLink: the witness crosses PCIe once.
        nr = info.num_routed_wires
        per = (nr + self.world - 1) // self.world
        lo, hi = min(nr, self.rank * per), min(nr, (self.rank + 1) * per)
        part = self.ops.empty(per * self.n).view(per, self.n)
        if hi > lo:
            ptrs = (C.c_void_p * (hi - lo))(*[wires_host[j].ctypes.data for j in range(lo, hi)])
            check(lib.eng_h2d_columns(ptrs, hi - lo, self.n, C.c_void_p(part.data_ptr())))
        routed = self._all_gather(part.view(-1)) if self.world > 1 else part
        zvals = self.ops.empty(self.nzs * self.n).view(self.nzs, self.n)
        check(lib.eng_partial_products_from_dev(self._h, C.c_void_p(routed.data_ptr()), _ptr(betas), _ptr(gammas), C.c_void_p(zvals.data_ptr())))
        del routed, part
        mark()
        zcols = self.plans[2].columns_of(self.rank)
        zs = self._commit(2, zvals[zcols.start:zcols.stop].contiguous(), True)
        del zvals
        ch.observe_cap(zs.cap)
        alphas = _u64(ch.get_n_challenges(self.nch))
        mark()
        # quotient: constraint evaluation on this rank's rows, all-gather, coset iNTT (replicated), commit of this rank's chunks
        qv = self.ops.empty(self.nch * self.count)
        check(lib.eng_quotient_values_shard_dev(self._h, C.c_void_p(self.cs.rows.data_ptr()), C.c_void_p(wires.rows.data_ptr()),
                                                C.c_void_p(zs.rows.data_ptr()), self.log_world, self.rank, _ptr(pi), _ptr(betas),
                                                _ptr(gammas), _ptr(alphas), C.c_void_p(qv.data_ptr())))
        gathered = self._all_gather(qv)
        qcoeffs = self.ops.empty(self.nq * self.n).view(self.nq, self.n)
        check(lib.eng_quotient_coeffs_from_shards_dev(self._h, C.c_void_p(gathered.data_ptr()), self.log_world, C.c_void_p(qcoeffs.data_ptr())))
        del gathered, qv
        mark()
        qcols = self.plans[3].columns_of(self.rank)
        quot = self._commit(3, qcoeffs[qcols.start:qcols.stop].contiguous(), False)
        del qcoeffs
        ch.observe_cap(quot.cap)
        zeta = ch.get_extension_challenge()
        g_n = pow(_POWER_OF_TWO_GENERATOR, 1 << (32 - info.degree_bits), P)
        gzeta = (zeta[0] * g_n % P, zeta[1] * g_n % P)
        mark()
        # opening set: every rank evaluates the polynomials of its column shards
        batches = [self.cs, wires, zs, quot]
        mine = [self._eval(b, zeta) for b in batches] + [self._eval(zs, gzeta)]
        parts = self._all_gather_object(mine)
        e_cs, e_w, e_z, e_q, e_zn = [np.concatenate([p[k] for p in parts]) for k in range(5)]
        ncs, nch = info.num_constants, self.nch
        op_consts, op_sig, op_zs, op_pp, op_zn = e_cs[:ncs], e_cs[ncs:], e_z[:nch], e_z[nch:], e_zn[:nch]
        for v in (op_consts, op_sig, e_w, op_zs, op_pp, e_q, op_zn):
            ch.observe_extension_elements(v.ravel())
        mark()
        # opening proofs: FRI layer 0 on this rank's rows, all-gather, the rest replicated; query openings from the row owners
        alpha = _u64(ch.get_extension_challenge())
        inst = self.fri_instance(zeta, gzeta).blob()
        openings = _u64(np.concatenate([e_cs, e_w, e_z, e_q, op_zn]))
        rows = (C.c_void_p * 4)(*[b.rows.data_ptr() for b in batches])
        widths = (C.c_uint32 * 4)(*self.widths)
        log_l = info.degree_bits + info.rate_bits
        layer = self.ops.empty(2 * self.count)
        check(lib.eng_fri_combine_shard_dev(_ptr(inst), rows, widths, 4, _ptr(openings), _ptr(alpha), log_l, self.log_world, self.rank,
                                            C.c_void_p(layer.data_ptr())))
        layer0 = self._all_gather(layer)
        params = FriParams(info.degree_bits, info.rate_bits, info.cap_height, self.pow_bits, self.num_query_rounds)
        blob = C.POINTER(C.c_uint64)()
        blen = C.c_size_t(0)
        x_idx = np.zeros(self.num_query_rounds, np.uint64)
        check(lib.eng_fri_prove_from_layer_dev(C.c_void_p(layer0.data_ptr()), ch._h, params.as_array(), C.byref(blob), C.byref(blen), _ptr(x_idx)))
        fri = np.ctypeslib.as_array(blob, shape=(blen.value,)).copy()
        lib.eng_blob_free(blob)
        del layer0, layer
        mine = {}
        for q, x in enumerate(int(v) for v in x_idx):
            if batches[0].owns_leaf(x):
                mine[q] = [(b.get(x), b.prove(x)) for b in batches]
        merged = {}
        for part in self._all_gather_object(mine):
            merged.update(part)
        fri = splice_initial_openings(fri, [merged[q] for q in range(self.num_query_rounds)])
        mark()
        proof = np.concatenate([_u64(wires.cap).ravel(), _u64(zs.cap).ravel(), _u64(quot.cap).ravel()] +
                               [_u64(v).ravel() for v in (op_consts, op_sig, e_w, op_zs, op_zn, op_pp, e_q)] + [fri])
        ms = [1e3 * (b_ - a_) for a_, b_ in zip(marks[:-1], marks[1:])]
        ms.append(1e3 * (marks[-1] - marks[0]))
        self.last_batches = batches      # keeps the exchange slots' batches alive for inspection
        return proof, dict(zip(self.STAGES, ms))

    def verify(self, public_inputs_hash, proof_blob):
        from .plonky2 import verify
        verify(self.blob, self.cs.cap, public_inputs_hash, proof_blob)

    def close(self):
        for ex in self.exchanges:
            if ex is not None:
                ex.close()
        self.exchanges = [None] * 4
        if getattr(self, "_h", None):
            _lib.lib().eng_circuit_free(self._h)
            self._h = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().eng_circuit_free(self._h)
                self._h = None
        except Exception:
            pass
