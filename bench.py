#!/usr/bin/env python3
"""bench.py -- commit-phase throughput of the B200 plonky2 engine (BASELINE.json configs[1]).

A step = one PolynomialBatch::from_values over synthetic witness columns: iNTT -> coset LDE (rate_bits 3) ->
Poseidon leaf hashing -> digest tree -> cap (cap_height 4).  Workload at N=1: 135 Goldilocks columns x 2^20 rows.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--log-n 20] [--cols 135]

metric  commit_throughput [GB/s] = algorithmic commit bytes B_ntt = 8*C*n*(2 + 2^rate_bits) per step / time
        (read the values once, write the coefficients once, write the LDE once).
value   inputs already resident in HBM, timed with CUDA events on the engine's stream.
e2e     the same commit through the reference-facing C-ABI call with HOST buffers (eng_batch_from_values with
        pinned host columns; H2D copies and the D2H read of the cap inside the timed region).
roofline       the dominant kernel (Poseidon leaf hashing): integer-pipe bound, 6,612 IMAD32 per permutation.
roofline_hbm   the NTT/LDE kernels against the measured HBM copy bandwidth.
cpu_baseline   the C++/OpenMP oracle (a restatement of plonky2's CPU algorithm -- the Rust prover itself cannot be
               built here) timed on the host cores on a bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RATE_BITS = 3
CAP_HEIGHT = 4
IMAD_PER_PERM = 6612
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` captures at configs[1]
# (profiles/r02_leaves_final.md, profiles/r01_ntt_v3.md); only quoted when the workload is that configuration
NCU_TRAFFIC_LEAVES = 9.086e9 + 0.311e9
NCU_TRAFFIC_NTT = (1.14 + 1.11 + 1.14 + 1.11 + 1.57 + 9.64 + 9.06 + 9.03) * 1e9
NOMINAL_IMAD_PER_S = 148 * 64 * 1.965e9   # 64 IMAD/clk/SM; no integer entry in MEASURED_PEAKS.json


def b_ntt(cols, n, rate_bits=RATE_BITS):
    return 8 * cols * n * (2 + (1 << rate_bits))


def num_perms(cols, n, rate_bits=RATE_BITS, cap_height=CAP_HEIGHT):
    L = n << rate_bits
    return L * ((cols + 7) // 8) + (L - (1 << cap_height))


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_commit(cols, log_n, threads=None):
    """Times the CPU oracle on one commit; returns (seconds, stage dict, threads)."""
    from oracle import oracle as O
    # all host cores, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1 to its workers)
    O.lib().orc_set_num_threads(threads or os.cpu_count() or 1)
    vals = O.splitmix_columns(cols, 1 << log_n)
    t0 = time.perf_counter()
    b = O.Batch.from_values(vals, RATE_BITS, CAP_HEIGHT)
    dt = time.perf_counter() - t0
    stages = b.times()
    del b
    return dt, stages, O.lib().orc_num_threads()


def cpu_sample_log_n(cols, target_s=12.0, max_log_n=20):
    """Largest sample (rows = 2^k <= 2^max_log_n) whose commit should take about target_s on this host.  ONE rule for
    the cpu_baseline object of the b200 arm and for every step of --impl reference (same target, same cap)."""
    k = 12
    dt, _, _ = cpu_commit(cols, k)
    while k < max_log_n:
        step = max(1, min(max_log_n - k, int(math.floor(math.log2(max(target_s / max(dt, 1e-6), 1.0)))) - 1))
        if dt * 2.0 > target_s:
            break
        k += step
        dt, _, _ = cpu_commit(cols, k)      # re-measure: small samples over-estimate the per-row cost
    while k > 12 and dt > 1.5 * target_s:
        k -= 1
        dt = dt / 2
    return k


def cpu_sample_desc(cols, k, dt, stages, threads):
    """The cpu_baseline object: value, the sample that was actually run, and the port's Poseidon rate per core."""
    perms = num_perms(cols, 1 << k)
    tree_s = stages.get("tree", 0.0)
    d = {"value": b_ntt(cols, 1 << k) / dt / 1e9, "unit": "GB/s", "cores": threads, "kind": "port", "sample_log_rows": k,
         "sample": "one PolynomialBatch::from_values of %d columns x 2^%d rows (rate_bits 3, cap_height 4), %.1f s; C++/OpenMP "
                   "restatement of plonky2's CPU algorithm (the Rust prover cannot be built here: no cargo, plonky2 not vendored)" % (cols, k, dt),
         "stage_s": stages}
    if tree_s > 0:
        d["poseidon_perms_per_s_per_core"] = perms / tree_s / max(1, threads)
        d["poseidon_us_per_perm_per_core"] = 1e6 * tree_s * max(1, threads) / perms
    return d


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the host cores, same metric / config / unit.
    The reference's own Rust prover cannot be built here (no cargo/rustc, plonky2 not vendored): the timed code is
    the oracle port (kind = "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cols = args.cols
    # every step is a bounded sample of the workload: the largest power-of-two row count (<= the workload's) whose commit
    # takes about 12 s on this host -- the same rule the b200 arm uses for its cpu_baseline object.  The row count that
    # was actually run is printed in config.sample_log_rows; value is a rate (GB/s of algorithmic commit bytes), which
    # is what makes a sample comparable with the full-size GPU step.
    k = min(args.log_n, cpu_sample_log_n(cols))
    for _ in range(args.warmup):
        cpu_commit(cols, k)
    t = []
    stages, threads = {}, 1
    for _ in range(args.steps):
        dt, stages, threads = cpu_commit(cols, k)
        t.append(dt)
    sec = sum(t) / len(t)
    gbs = b_ntt(cols, 1 << k) / sec / 1e9
    cfg = workload_config(cols, args.log_n, args.gpus)
    cfg["sample_log_rows"] = k
    cfg["sample_note"] = ("CPU arm: each step commits %d columns x 2^%d rows (a bounded sample of the 2^%d-row workload, about %.0f s per "
                          "step on %d threads); rates, not times, are comparable" % (cols, k, args.log_n, sec, threads))
    line = {
        "impl": "reference", "metric": "commit_throughput", "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 (Goldilocks field, integer)", "data": "synthetic (SplitMix64 witness columns)",
        "config": cfg,
        "cpu_baseline": cpu_sample_desc(cols, k, sec, stages, threads),
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def oracle_verify(blob, cs_cap, pi_hash, proof):
    """Checker use of the oracle (allowed outside the timed region): its restatement of plonky2's verifier on a bench proof."""
    try:
        from oracle import oracle as O
        return O.Circuit(blob).verify(cs_cap, pi_hash, proof) == 0
    except Exception as ex:   # noqa: BLE001
        return "unavailable: %r" % (ex,)


def proof_section(E, log_n, reps=4, all_gates=False):
    """BASELINE.json's first metric: proof wall-time (s).  The real eth-lc circuit cannot be built here (it needs
    plonky2's Rust front-end), so this is the circuit-SHAPED synthetic proof of SURVEY.md 8(d): 135 wires, 80 routed,
    selectors + 2 constants + 80 sigmas, standard_recursion_config (rate_bits 3, cap_height 4, 16-bit grind, 28 queries),
    witness on the host, everything after witness generation on the GPU through eng_prove.  Gate set: the five core gates
    (half the rows PoseidonGate), or with all_gates the 22 gate kinds of gate_lib.h (recursion + plonky2_crypto u32 gates,
    evaluated from bytecode) -- the gate mix of the real circuit.  Wall clock around the call with host wire columns (H2D
    inside).  Every proof is verified (product verifier + the oracle's restatement of plonky2's)."""
    import hashlib
    try:      # start from an empty pool: a long-lived process that has run other shapes has a fragmented one
        E.release_cached()
    except Exception:   # noqa: BLE001
        pass
    s = E.synth_circuit_v2(log_n, seed=1) if all_gates else E.synth_circuit(log_n, seed=1)
    pinned = True
    try:      # the witness lives in page-locked host memory (eng_host_register), as the e2e contract assumes
        E.host_register(s["wires"])
    except Exception:   # noqa: BLE001
        pinned = False
    t0 = time.perf_counter()
    circ = E.Circuit.build(s)
    E.synchronize()
    build_s = time.perf_counter() - t0
    wires = list(s["wires"])
    walls, stages, best = [], {}, None
    for i in range(reps):
        t = time.perf_counter()
        proof, st = circ.prove(wires, s["pi_hash"])
        walls.append(time.perf_counter() - t)
        if i and (best is None or walls[-1] < best):     # the first call pays whatever the process has not warmed yet
            best, stages = walls[-1], st
    if best is None:
        best, stages = walls[0], st
    verified = {"engine": False, "oracle": None}
    try:
        circ.verify(s["pi_hash"], proof)
        verified["engine"] = True
    except Exception as ex:   # noqa: BLE001
        verified["engine"] = repr(ex)[:200]
    verified["oracle"] = oracle_verify(s["blob"], circ.constants_sigmas.merkle_tree.cap, s["pi_hash"], proof)
    if pinned:
        E.host_unregister(s["wires"])
    return {"metric": "synthetic_proof_wall_time", "value": best, "unit": "s", "all_runs_s": walls, "witness_memory": "page-locked (eng_host_register)" if pinned else "pageable",
            "higher_is_better": False, "log_rows": log_n,
            "config": "circuit-shaped synthetic proof, 2^%d rows x 135 wires, %s, standard_recursion_config" % (
                log_n, "22 gate kinds (core + recursion + u32; bytecode description, library gates compiled)" if all_gates else "5 core gates"),
            "build_constants_sigmas_commit_s": build_s, "first_call_in_this_process_s": build_s + walls[0],
            "first_call_note": "Circuit.build (constants||sigmas commit + memory-pool growth for one proof) + the first eng_prove of this "
                               "circuit inside the bench process; the cold start of a FRESH process is the line's cold_start object",
            "stage_ms": stages, "proof_u64_words": int(proof.size), "proof_sha256": hashlib.sha256(proof.tobytes()).hexdigest(),
            "verified": verified,
            "reference_published": "~300 s for the real 2^22-row circuit on 32 vCPU (README.md:71); not comparable 1:1"}


def sharded_proof_section(E, log_n, rank, world, local_rank, reps=3, compare_single=True):
    """N > 1: the same synthetic proof through ShardedProver (parallel.py): strong scaling of ONE proof over the ranks.
    Every rank builds the same synthetic circuit and witness (seeded).  The proof is verified and, on rank 0, compared
    word for word with the single-GPU eng_prove of the same witness."""
    import hashlib
    import numpy as np
    import torch
    import torch.distributed as dist
    s = E.synth_circuit(log_n, seed=1)
    pinned = True
    try:
        E.host_register(s["wires"])
    except Exception:   # noqa: BLE001
        pinned = False
    t0 = time.perf_counter()
    pr = E.ShardedProver(s["blob"], s["constants"], s["sigmas"], rank, world, device=torch.device("cuda", local_rank))
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    walls, stages = [], {}
    for i in range(reps):
        dist.barrier()
        torch.cuda.synchronize()
        t = time.perf_counter()
        proof, st = pr.prove(s["wires"], s["pi_hash"])
        walls.append(time.perf_counter() - t)
        if i == reps - 1:
            stages = st
    tw = torch.tensor(walls, dtype=torch.float64, device="cuda")
    dist.all_reduce(tw, op=dist.ReduceOp.MAX)          # a proof is done when the slowest rank has it
    walls = tw.tolist()
    sha = hashlib.sha256(proof.tobytes()).hexdigest()
    shas = [None] * world
    dist.all_gather_object(shas, sha)
    out = None
    if rank == 0:
        verified = {"engine": False, "oracle": None, "identical_on_all_ranks": len(set(shas)) == 1}
        try:
            pr.verify(s["pi_hash"], proof)
            verified["engine"] = True
        except Exception as ex:   # noqa: BLE001
            verified["engine"] = repr(ex)[:200]
        verified["oracle"] = oracle_verify(s["blob"], pr.cs.cap, s["pi_hash"], proof)
        out = {"metric": "synthetic_proof_wall_time", "value": min(walls[1:]) if len(walls) > 1 else walls[0], "unit": "s", "all_runs_s": walls,
               "higher_is_better": False, "log_rows": log_n, "n_gpus": world, "scaling": "strong",
               "config": "circuit-shaped synthetic proof, 2^%d rows x 135 wires, 5 core gates, standard_recursion_config; ONE proof "
                         "over %d GPUs (ShardedProver: column/row-sharded commits, row-local quotient and FRI layer 0)" % (log_n, world),
               "build_constants_sigmas_commit_s": build_s, "stage_ms": stages, "proof_sha256": sha, "verified": verified}
    pr.close()
    if pinned:
        E.host_unregister(s["wires"])
    if rank == 0 and compare_single:
        try:
            circ = E.Circuit.build(s)
            ref, _ = circ.prove(list(s["wires"]), s["pi_hash"])
            out["verified"]["equals_single_gpu_proof"] = bool(ref.shape == proof.shape and (ref == proof).all())
            del circ
            E.release_cached()
        except Exception as ex:   # noqa: BLE001
            out["verified"]["equals_single_gpu_proof"] = "not compared: %r" % (ex,)
    dist.barrier()
    return out


def peer_parity_check(E, plan_cols, rank, world, local_rank, use_peer):
    """N > 1, before timing: a 135 x 2^14 commit through the REAL exchange (CUDA-IPC peer stores, or NCCL all-to-all) against
    (a) the single-GPU engine path and (b) the oracle, on rank 0: cap, every digest, Merkle paths (VERDICT r1 task 1b)."""
    import hashlib
    import numpy as np
    import torch
    import torch.distributed as dist
    log_n = 14
    n = 1 << log_n
    plan = E.ShardPlan(plan_cols, log_n, RATE_BITS, CAP_HEIGHT, world)
    ex = None
    if use_peer:
        ex = E.PeerExchange(plan, rank, torch.device("cuda", local_rank))
    cols = plan.columns_of(rank)
    vals = E.splitmix_columns(len(cols), n, first_col=cols.start)
    dev = torch.from_numpy(vals.view(np.int64)).cuda()
    torch.cuda.synchronize()
    b = E.ShardedPolynomialBatch.from_values(dev, plan, rank, exchange=ex)
    digs = [None] * world
    dist.all_gather_object(digs, hashlib.sha256(np.ascontiguousarray(b.local_digests).tobytes()).hexdigest())
    L = n << RATE_BITS
    probe = [0, 1, L // 2 + 77, L - 1] + [r * plan.rows_per_rank + 5 for r in range(world)]
    mine = {k: (b.get(k).tolist(), b.prove(k).tolist()) for k in probe if b.owns_leaf(k)}
    opened = [None] * world
    dist.all_gather_object(opened, mine)
    res = None
    if rank == 0:
        full = E.splitmix_columns(plan_cols, n)
        single = E.PolynomialBatch.from_values(list(full), RATE_BITS, False, CAP_HEIGHT)
        sd = single.merkle_tree.digests
        per = sd.shape[0] // world
        res = {"shape": "%d columns x 2^%d rows, rate_bits 3, cap_height 4, %d ranks" % (plan_cols, log_n, world),
               "exchange": "peer stores (CUDA IPC)" if ex is not None else "NCCL all_to_all",
               "cap_equals_single_gpu": bool((np.array(b.cap) == single.merkle_tree.cap).all()),
               "digests_equal_single_gpu": all(digs[r] == hashlib.sha256(np.ascontiguousarray(sd[r * per:(r + 1) * per]).tobytes()).hexdigest() for r in range(world)),
               "cap_sha256": hashlib.sha256(np.ascontiguousarray(b.cap).tobytes()).hexdigest()}
        ok = True
        for part in opened:
            for k, (leaf, path) in part.items():
                ok = ok and leaf == single.merkle_tree.get(k).tolist() and path == single.merkle_tree.prove(k).tolist()
        res["leaves_and_paths_equal_single_gpu"] = bool(ok)
        try:
            from oracle import oracle as O
            o = O.Batch.from_values(full, RATE_BITS, CAP_HEIGHT)
            res["cap_equals_oracle"] = bool((np.array(b.cap) == o.cap).all())
            res["digests_equal_oracle"] = bool((sd == o.digests).all())
            res["paths_equal_oracle"] = all(path == o.prove(k).tolist() for part in opened for k, (leaf, path) in part.items())
        except Exception as exn:   # noqa: BLE001
            res["oracle"] = "unavailable: %r" % (exn,)
        res["peer_exchange"] = "ok" if all(v for v in res.values() if isinstance(v, bool)) else "MISMATCH"
        single.close()
    del b
    if ex is not None:
        ex.close()
    dist.barrier()
    return res


def cold_start_section(log_n):
    """Cold start of a prover process (VERDICT r1 task 6): tools/prof_coldstart.py in a FRESH process -- import, eng_init,
    Circuit.build (constants||sigmas commit + pool growth) and three proofs of the 2^log_n-row synthetic circuit."""
    import ast
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "prof_coldstart.py"), str(log_n), "v1"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=600)
    line = [l for l in res.stdout.splitlines() if l.startswith("{")]
    if res.returncode != 0 or not line:
        return {"error": "prof_coldstart.py exited %d" % res.returncode}
    d = ast.literal_eval(line[-1])
    return {"log_rows": log_n, "eng_init_s": d["eng_init_s"], "circuit_build_s": d["build_s"], "first_proof_s": d["prove_s"][0],
            "second_proof_s": d["prove_s"][1], "first_call_s": d["build_s"] + d["prove_s"][0], "verified": d["verified"],
            "note": "fresh process; the memory pool is grown once in Circuit.build (eng_reserve), so the first proof costs what every proof costs; "
                    "the growth rate (20-60 GB/s) is the driver's and varies from box to box"}


def dist_roofline(stages, cols, n, world):
    """N > 1: rank 0's leaf hashing (the dominant kernel, identical work on every rank) against the integer-pipe roofline."""
    try:
        leaf_ms = stages.get("build Merkle tree (leaves)", 0.0)
        if leaf_ms <= 0:
            return None
        perms = ((n << RATE_BITS) // world) * ((cols + 7) // 8)      # per rank and launch
        rate = perms * IMAD_PER_PERM / (leaf_ms * 1e-3)
        return {"kernel": "merkle_leaves_kernel on rank 0 (%d permutations per launch)" % perms, "bound": "int32-imad",
                "achieved": rate / 1e12, "peak": NOMINAL_IMAD_PER_S / 1e12, "unit": "TIMAD32/s", "frac": rate / NOMINAL_IMAD_PER_S,
                "traffic": None, "kernel_ms": leaf_ms}
    except Exception:   # noqa: BLE001
        return None


def exchange_close(exchange):
    if exchange is not None:
        exchange.close()


def workload_config(cols, log_n, gpus, scaling="weak"):
    return {"workload": "PolynomialBatch::from_values commit, %d Goldilocks columns x 2^%d rows, rate_bits=3, cap_height=4 "
                        "(BASELINE.json configs[1])" % (cols, log_n),
            "columns": cols, "log_rows": log_n, "rate_bits": RATE_BITS, "cap_height": CAP_HEIGHT,
            "parallelism": "1 GPU" if gpus == 1 else "%d GPUs: column-sharded iNTT/LDE -> all-to-all -> row-sharded hashing -> cap "
                                                      "all-gather; %s" % (gpus, "rows scale with N (2^20 per GPU)" if scaling == "weak" else "rows fixed (strong scaling)"),
            "l2": "inputs (%.2f GB per step) are larger than the 126 MB L2; no flush needed" % (8 * cols * (1 << log_n) / 1e9)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--cols", type=int, default=135)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--proof-log-n", type=int, default=20, help="rows (log2) of the synthetic circuit-shaped proof; 0 = skip")
    ap.add_argument("--proof-full-log-n", type=int, default=22, help="rows (log2) of the full-size synthetic proof; 0 = skip")
    ap.add_argument("--dist-proof-log-n", type=int, default=20, help="N > 1: rows (log2) of the ONE synthetic proof run over all ranks (strong scaling); 0 = skip")
    ap.add_argument("--dist-proof-full-log-n", type=int, default=22, help="N > 1: the eth-lc-sized proof (2^22 rows) over all ranks; 0 = skip")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1 commit: weak = 2^log_n rows PER GPU (default, the driver's scaling run); strong = 2^log_n rows in total")
    ap.add_argument("--all-gates-proof-log-n", type=int, default=20, help="N = 1: rows (log2) of the synthetic proof over all 22 gate kinds; 0 = skip")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: column->row exchange fused into the LDE's last pass as NVLink peer stores (default), or NCCL all-to-all")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import eth_lc_plonky2_b200 as E

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    distributed = world > 1
    torch.cuda.set_device(local_rank)
    if distributed:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    E.init(local_rank)
    stream = torch.cuda.Stream()
    E.set_stream(stream.cuda_stream)

    cols = args.cols
    # N = 1: BASELINE configs[1].  N > 1: the same commit column-sharded over the ranks with one all-to-all
    # (parallel.py); rows scale with N so that the per-GPU work is fixed (weak scaling, configs[4] sweep shape).
    log_n = args.log_n + (world.bit_length() - 1 if args.scaling == "weak" else 0)
    n = 1 << log_n
    parity = None
    if distributed:      # real-transport parity before anything is timed
        try:
            parity = peer_parity_check(E, cols, rank, world, local_rank, args.exchange == "peer")
        except Exception as ex:   # noqa: BLE001
            parity = {"peer_exchange": "check failed to run: %r" % (ex,)}
    plan = E.ShardPlan(cols, log_n, RATE_BITS, CAP_HEIGHT, world) if distributed else None
    exchange = None
    if distributed and args.exchange == "peer":
        try:      # collective: either every rank gets the fused exchange or every rank falls back to the NCCL all-to-all
            exchange = E.PeerExchange(plan, rank, torch.device("cuda", local_rank))
        except E.EngineError as ex:
            if rank == 0:
                print("bench: %s -- falling back to --exchange nccl" % ex, file=sys.stderr)
    my_cols = plan.columns_of(rank) if distributed else range(cols)
    # synthetic witness, generated on the host, pinned for the e2e path
    host = torch.from_numpy(E.splitmix_columns(len(my_cols), n, first_col=my_cols.start).view(np.int64)).pin_memory()
    dev = host.cuda(non_blocking=False)
    host_np = host.numpy().view(np.uint64)
    host_cols = [host_np[c] for c in range(len(my_cols))]
    stage_keys = ("IFFT", "FFT + blinding", "transpose LDEs", "build Merkle tree (leaves)", "build Merkle tree (digest levels)", "host to device")

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        if distributed:
            b = E.ShardedPolynomialBatch.from_values(dev, plan, rank, exchange=exchange)
            ms = {k: 0.0 for k in stage_keys}
            try:     # this rank's hashing times (CUDA events inside eng_merkle_new_dev); the transform is timed as a whole
                t = (ctypes.c_float * 6)()
                if E._lib.lib().eng_batch_stage_ms(b.merkle_tree_local._o._h, t) == 0:
                    ms["build Merkle tree (leaves)"], ms["build Merkle tree (digest levels)"] = float(t[3]), float(t[4])
            except Exception:   # noqa: BLE001
                pass
            return ms
        b = E.PolynomialBatch.from_values(dev, RATE_BITS, False, CAP_HEIGHT)
        ms = b.stage_ms()
        b.close()
        return ms

    def step_e2e():
        if distributed:
            if exchange is not None:                       # host columns straight into the copy / transform / peer-store pipeline
                b = E.ShardedPolynomialBatch.from_values(host_cols, plan, rank, exchange=exchange)
            else:
                staged = host.cuda(non_blocking=True)      # H2D of this rank's columns
                b = E.ShardedPolynomialBatch.from_values(staged, plan, rank)
            return b.cap                                   # replicated cap, already on the host
        b = E.PolynomialBatch.from_values(host_cols, RATE_BITS, False, CAP_HEIGHT)
        cap = b.merkle_tree.cap            # D2H read of the result
        b.close()
        return cap

    # ---- device-resident arm ----
    # warm-up on the SAME torch stream as the timed steps: torch's caching allocator keeps its blocks per stream, so a
    # warm-up on the default stream would leave the first timed step to allocate its multi-GB scratch afresh (measured at
    # N = 2: 460 ms for the first timed step against 147 ms for the others)
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    l0 = E.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_acc = {}
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            t_step = time.perf_counter()
            for k, v in step_device().items():
                stage_acc[k] = stage_acc.get(k, 0.0) + v
            if os.environ.get("ENG_TRACE") and rank == 0:
                print("rank 0: device step wall %.1f ms" % (1e3 * (time.perf_counter() - t_step)), flush=True)
        ev1.record(stream)
    barrier()
    launches = E.launch_count() - l0
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.summary()

    # ---- end-to-end arm (host buffers through the C ABI) ----
    with torch.cuda.stream(stream):
        for _ in range(max(1, min(args.warmup, 3))):
            cap_e2e = step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            cap_e2e = step_e2e()
        e1.record(stream)
    barrier()
    ms_e2e = e0.elapsed_time(e1)

    times = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = times.tolist()

    int_peak = E.measure_int_peak()
    import hashlib
    cap_sha = hashlib.sha256(np.ascontiguousarray(cap_e2e).tobytes()).hexdigest()
    dist_proofs = {}
    if distributed:
        # release the commit bench's buffers before the proofs (every rank; the exchange of the commit stays open until the end)
        try:
            del dev
            torch.cuda.empty_cache()
            E.release_cached()
        except Exception:   # noqa: BLE001
            pass
        for key, lg in (("proof", args.dist_proof_log_n), ("proof_full_size", args.dist_proof_full_log_n)):
            if not lg:
                continue
            try:     # collective: every rank takes part; a failure on any rank must not lose the commit line
                dist_proofs[key] = sharded_proof_section(E, lg, rank, world, local_rank)
            except Exception as ex:   # noqa: BLE001
                dist_proofs[key] = {"error": repr(ex)[:300]}
                break
    if rank == 0:
        peaks, peak_kind = measured_peaks()
        bytes_step = b_ntt(cols, n)
        ms_step = ms_total / args.steps
        value = bytes_step / (ms_step * 1e-3) / 1e9
        e2e_value = bytes_step / (ms_e2e / args.steps * 1e-3) / 1e9
        stages = {k: v / args.steps for k, v in stage_acc.items()}
        if distributed:
            line = {
                "metric": "commit_throughput", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "u64 (Goldilocks field, integer)", "data": "synthetic (SplitMix64 witness columns)",
                "config": workload_config(cols, log_n, world, args.scaling),
                "e2e": {"value": e2e_value, "unit": "GB/s", "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": 8 * cols * n, "d2h_bytes_per_step": (32 << CAP_HEIGHT) * world},
                "gpu_launches": launches, "clocks": clocks, "cap_sha256": cap_sha, "parity_check": parity,
                "stage_ms": stages, "roofline": dist_roofline(stages, cols, n, world),
                "exchange": ("fused: the LDE's last pass stores %.2f GB per rank into the peers' leaf matrices (NVLink P2P, CUDA IPC); NCCL carries two "
                             "barriers and the cap all_gather" if exchange is not None else
                             "all_to_all_single of %.2f GB per rank (NCCL), cap all_gather") % (8 * len(my_cols) * (n << RATE_BITS) * (world - 1) / world / 1e9),
            }
            line.update({k: v for k, v in dist_proofs.items() if v is not None})
            print(json.dumps(line), flush=True)
            exchange_close(exchange)
            dist.destroy_process_group()
            return
        leaf_ms = stages["build Merkle tree (leaves)"]
        L = n << RATE_BITS
        leaf_perms = L * ((cols + 7) // 8)
        imad_rate = leaf_perms * IMAD_PER_PERM / (leaf_ms * 1e-3)
        ntt_ms = stages["IFFT"] + stages["FFT + blinding"]
        hbm_rate = b_ntt(cols, n) / (ntt_ms * 1e-3) / 1e9
        line = {
            "metric": "commit_throughput", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64 (Goldilocks field, integer)", "data": "synthetic (SplitMix64 witness columns)",
            "config": workload_config(cols, log_n, world),
            "e2e": {"value": e2e_value, "unit": "GB/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": 8 * cols * n, "d2h_bytes_per_step": 32 << CAP_HEIGHT},
            "gpu_launches": launches,
            "stage_ms": stages,
            "roofline": {"kernel": "merkle_leaves_kernel (Poseidon leaf hashing, %d permutations per launch)" % leaf_perms,
                         "bound": "int32-imad", "achieved": imad_rate / 1e12, "peak": NOMINAL_IMAD_PER_S / 1e12,
                         "unit": "TIMAD32/s", "frac": imad_rate / NOMINAL_IMAD_PER_S,
                         "traffic": NCU_TRAFFIC_LEAVES if (cols, log_n) == (135, 20) else None,
                         "hbm_floor_ms": (8 * cols * L + 32 * L) / (peaks["hbm_gbs"] * 1e9) * 1e3,
                         # the same kernel against the HBM roof (it is not memory bound: one leaf is 1080 B and 17 permutations)
                         "hbm": {"bound": "hbm", "achieved": (8 * cols * L + 32 * L) / (leaf_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                                 "unit": "GB/s", "frac": (8 * cols * L + 32 * L) / (leaf_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]},
                         "peak_kind": "nominal 148 SM x 64 IMAD/clk x 1.965 GHz (no measured integer peak in MEASURED_PEAKS.json)",
                         "perms_per_s": leaf_perms / (leaf_ms * 1e-3), "kernel_ms": leaf_ms,
                         "measured_issue_rates_Tops": {k: v / 1e12 for k, v in int_peak.items()},
                         "frac_of_measured_imad": imad_rate / int_peak["imad"]},
            "roofline_hbm": {"kernel": "ntt_pass_kernel x4 (iNTT 2 passes + coset LDE 2 passes)", "bound": "hbm",
                             "achieved": hbm_rate, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_rate / peaks["hbm_gbs"],
                             "traffic": NCU_TRAFFIC_NTT if (cols, log_n) == (135, 20) else None, "peak_kind": peak_kind,
                             "note": "integer-issue bound, not HBM bound: alu pipe 57-69 % of peak at 54-61 % issue activity, DRAM 7-16 % "
                                     "(profiles/r01_ntt_v3.md); reported against the HBM roof because BASELINE.json asks for it", "kernel_ms": ntt_ms, "algorithmic_bytes": b_ntt(cols, n)},
            "clocks": clocks,
            "cap_sha256": cap_sha,
        }
        # the proof sections are extras next to the headline metric: a failure there must not lose the commit line
        if args.proof_log_n:
            try:
                line["proof"] = proof_section(E, args.proof_log_n)
            except Exception as ex:   # noqa: BLE001
                line["proof"] = {"error": repr(ex)[:300]}
        if args.all_gates_proof_log_n:
            # the gate mix of the real circuit: 22 gate kinds evaluated from bytecode (gate_vm.h)
            try:
                line["proof_all_gates"] = proof_section(E, args.all_gates_proof_log_n, reps=3, all_gates=True)
            except Exception as ex:   # noqa: BLE001
                line["proof_all_gates"] = {"error": repr(ex)[:300]}
        if args.proof_full_log_n:
            # the eth-lc circuit's own size (~2.98 M constraints => 2^22 rows, BASELINE configs[3]); same synthetic gate set
            try:
                line["proof_full_size"] = proof_section(E, args.proof_full_log_n, reps=3)
            except Exception as ex:   # noqa: BLE001
                line["proof_full_size"] = {"error": repr(ex)[:300]}
        if args.proof_full_log_n and world == 1:
            try:
                E.release_cached()
                torch.cuda.empty_cache()
                line["cold_start"] = cold_start_section(args.proof_full_log_n)
            except Exception as ex:   # noqa: BLE001
                line["cold_start"] = {"error": repr(ex)[:300]}
        if not args.no_cpu_baseline and world == 1:
            try:
                E.release_cached()    # give the device memory of the proof sections back before the host-side baseline
            except Exception:         # noqa: BLE001
                pass
            k = min(args.log_n, cpu_sample_log_n(cols))
            dt, cpu_stages, threads = cpu_commit(cols, k)
            line["cpu_baseline"] = cpu_sample_desc(cols, k, dt, cpu_stages, threads)
        print(json.dumps(line), flush=True)
    if distributed:
        exchange_close(exchange)      # collective: every rank
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
