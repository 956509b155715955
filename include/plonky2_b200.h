/* plonky2_b200.h -- C ABI of the B200-native plonky2 commitment engine (libplonky2_b200.so).
 *
 * Drop-in boundary for the hot path of Electron-Labs/eth-lc-plonky2: the operators of the un-vendored
 * dependency plonky2 0.1.4 (git 666f3151..., pinned at /root/reference/Cargo.lock:2347-2350) that
 * `builder.build::<C>()` and `data.prove(witness)` spend their time in
 * (/root/reference/eth-lc-plonky2/src/main.rs:227 and :230; every test goes through
 * /root/reference/eth-lc-plonky2/src/unit_tests.rs:29-35).  plonky2 has no plugin trait for them; the seam is a
 * Cargo [patch] of the git dependency (/root/reference/Cargo.toml:21-23,
 * /root/reference/eth-lc-plonky2/Cargo.toml:9) whose operator bodies call these symbols -- see INTEGRATION.md.
 *
 * Conventions
 *   - Field elements are little-endian uint64_t (GoldilocksField is #[repr(transparent)] over u64); inputs may be
 *     non-canonical, every output is the canonical representative (< p = 2^64 - 2^32 + 1).
 *   - Digests (HashOut<F>) are 4 x uint64_t.
 *   - Every function returns an eng_status; it never unwinds.  On failure eng_last_error() gives the message that
 *     the Rust shim turns into the panic!/anyhow! plonky2 would have raised
 *     (the four #[should_panic] tests, /root/reference/eth-lc-plonky2/src/unit_tests.rs:377,555,654,686).
 *   - Pointers named *_host are host memory owned by the caller for the duration of the call; *_dev are device
 *     pointers on the engine's device.  Batches are device-resident behind opaque handles (Rust: Drop ->
 *     eng_batch_free).  The API is re-entrant (one internal lock; `cargo test` runs proofs on parallel threads).
 *   - There is no CPU path: without a CUDA device eng_init fails and every other call returns ENG_ERR_STATE.
 */
#ifndef PLONKY2_B200_H
#define PLONKY2_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t eng_status;
#define ENG_OK 0
#define ENG_ERR_INVALID 1 /* a plonky2 assert!/expect would have fired (bad sizes, cap_height > log2(leaves), ...) */
#define ENG_ERR_CUDA 2    /* CUDA runtime error (sticky errors are reported on every later call) */
#define ENG_ERR_OOM 3     /* device allocation failed */
#define ENG_ERR_STATE 4   /* engine not initialised / no device */

typedef struct eng_batch eng_batch; /* PolynomialBatch (or a bare MerkleTree) resident on the device */

/* ---- lifecycle ---- */
eng_status eng_init(int32_t device);             /* idempotent; device = CUDA ordinal (one process per GPU) */
eng_status eng_shutdown(void);
eng_status eng_last_error(char *buf, size_t len); /* message of the calling thread's last failure */
eng_status eng_set_stream(void *cuda_stream);     /* run on the caller's cudaStream_t (NULL = engine's own) */
eng_status eng_synchronize(void);
/* Released device buffers are kept (by exact size) for the next call of the same shape; this returns them, and the
 * stream-ordered pool's unused memory, to the driver. */
eng_status eng_release_cached(void);
/* Grows the device memory pool so that `bytes` are free in it (cold start: the first proof of a process otherwise pays
 * the pool's growth buffer by buffer).  eng_circuit_new / eng_circuit_load call it with the footprint of one eng_prove
 * (option "reserve_for_proof", default 1); a prover can also call it early, while the host still builds the circuit. */
eng_status eng_reserve(size_t bytes);
/* Page-lock / release a caller-owned host range (witness columns): host-column entry points then copy at PCIe speed
 * straight from it instead of through the engine's pinned bounce buffers. */
eng_status eng_host_register(const void *ptr, size_t bytes);
eng_status eng_host_unregister(const void *ptr);
eng_status eng_launch_count(uint64_t *out);       /* kernels launched by the engine since eng_init */
/* Engine options (A/B switches used by tests and profiles; defaults are the product path):
 *   "quot_native_gates"    1  library gates whose program is gate_lib.h's run through the compiled evaluators (the same source
 *                             instantiated for the device) instead of the bytecode interpreter (0: interpret everything)
 *   "quot_native_poseidon" 1  PoseidonGate through the native FP64 evaluator (0: through its bytecode, like every other gate)
 *   "lde_group_mb"         0  megabytes of four-step intermediate a column group may hold between the two passes of a
 *                             transform so that the second pass reads it from the 126 MB L2 instead of HBM
 *                             (0: one launch per pass over the whole batch; measured in profiles/r02_ntt.md)
 *   "lde_peer_chunk_cols"  4  eng_lde_peer_dev: iNTT + LDE alternate over chunks of this many columns, which spreads the peer
 *                             stores over the whole transform (8 GPUs, 17 columns x 2^23 rows: 47.4 ms at once, 43.6 ms in
 *                             chunks of 4; 0 = the whole column shard at once)
 *   "reserve_for_proof"    1  eng_circuit_new / eng_circuit_load grow the memory pool to one proof's footprint */
eng_status eng_set_option(const char *name, int64_t value);

/* Measured integer issue rates of this device, thread-operations per second:
 * [0] 32-bit IMAD (mad.lo.u32)  [1] IMAD.WIDE (mad.wide.u32)  [2] alu-pipe ops (add/xor).  The roofline denominators
 * for the Poseidon kernels (MEASURED_PEAKS.json only has HBM and bf16). */
eng_status eng_measure_int_peak(double ops_per_s[3]);

/* ---- a4: PoseidonHash / Poseidon::poseidon  [plonky2:hash/poseidon.rs, hash/hashing.rs] ---- */
/* Poseidon::poseidon on `count` states of 12 lanes. */
eng_status eng_poseidon_permute(const uint64_t *states_host, uint64_t *out_host, size_t count);
/* PoseidonHash::hash_no_pad (or_noop = 0) / hash_or_noop (or_noop = 1) of `count` inputs of `len` elements each. */
eng_status eng_hash_n(const uint64_t *in_host, size_t len, size_t count, int32_t or_noop, uint64_t *out_host);
/* PoseidonHash::two_to_one on `count` pairs [l(4) | r(4)]. */
eng_status eng_two_to_one(const uint64_t *pairs_host, size_t count, uint64_t *out_host);

/* ---- a1/a2: PolynomialBatch::from_values / from_coeffs  [plonky2:fri/oracle.rs] ----
 * cols_host[c] points at polynomial c (2^log_n elements).  blinding != 0 appends SALT_SIZE = 4 random leaf
 * elements per row.  plonky2 draws them from OsRng; here they are the ChaCha20 key stream under a fresh 256-bit key taken
 * from the OS random number generator (getrandom) for every batch: blinding_seed MUST be 0 for that.  A non-zero
 * blinding_seed derives the key from the seed instead -- reproducible, for tests only, NOT hiding.
 * timing / fft_root_table of the Rust signature have no counterpart: stage times are read back with
 * eng_batch_stage_ms, root tables are cached inside the engine. */
eng_status eng_batch_from_values(const uint64_t *const *cols_host, uint32_t num_polys, uint32_t log_n,
                                 uint32_t rate_bits, int32_t blinding, uint64_t blinding_seed, uint32_t cap_height,
                                 eng_batch **out);
eng_status eng_batch_from_coeffs(const uint64_t *const *cols_host, uint32_t num_polys, uint32_t log_n,
                                 uint32_t rate_bits, int32_t blinding, uint64_t blinding_seed, uint32_t cap_height,
                                 eng_batch **out);
/* Same, input already on the device as [num_polys][2^log_n] (column-major, contiguous). */
eng_status eng_batch_from_values_dev(const uint64_t *values_dev, uint32_t num_polys, uint32_t log_n, uint32_t rate_bits,
                                     int32_t blinding, uint64_t blinding_seed, uint32_t cap_height, eng_batch **out);
eng_status eng_batch_from_coeffs_dev(const uint64_t *coeffs_dev, uint32_t num_polys, uint32_t log_n, uint32_t rate_bits,
                                     int32_t blinding, uint64_t blinding_seed, uint32_t cap_height, eng_batch **out);
eng_status eng_batch_free(eng_batch *b);

/* Multi-GPU building block: iNTT (is_values) + coset LDE of this rank's column shard into
 * caller-owned device buffers, asynchronously on the engine's stream.  coeffs_out_dev is [num_polys][n];
 * lde_out_dev is [G][num_polys][L/G] with G = 2^log_row_shards, L = n * 2^rate_bits: slice g holds LDE rows
 * [g*L/G, (g+1)*L/G) (bit-reversed order) of every local column, i.e. the contiguous all-to-all send chunk for
 * row-shard owner g.  src_dev may alias coeffs_out_dev.  The receiver hashes its rows with eng_merkle_new_dev. */
eng_status eng_lde_dev(const uint64_t *src_dev, uint32_t num_polys, uint32_t log_n, uint32_t rate_bits, int32_t is_values,
                       uint32_t log_row_shards, uint64_t *coeffs_out_dev, uint64_t *lde_out_dev);

/* Fused exchange: the same transforms, but the LAST pass of the LDE stores row shard g through
 * shard_out[g] -- a device pointer, normally a peer mapping (NVLink P2P) into row-shard owner g's leaf matrix
 * [C][L/G] at this rank's first column -- so the column->row all-to-all is the store itself and no send buffer is
 * re-read.  scratch_dev ([num_polys][L], local) holds the first pass' intermediate.  The caller synchronises all ranks
 * (stream sync + barrier) before hashing the received rows.  log_row_shards <= 4.  first_shard = the caller's own rank: the
 * last pass walks the row shards starting there, so that at any moment the ranks store into different destinations (an
 * all-to-all schedule) instead of all into shard 0, then 1, ... (measured at 8 GPUs: 56.7 ms without). */
eng_status eng_lde_peer_dev(const uint64_t *src_dev, uint32_t num_polys, uint32_t log_n, uint32_t rate_bits, int32_t is_values,
                            uint32_t log_row_shards, uint64_t *coeffs_out_dev, uint64_t *scratch_dev, uint64_t *const *shard_out,
                            uint32_t first_shard);
/* Same from HOST columns (pinned or pageable): column chunks are copied on a second stream while the previous chunk runs
 * its iNTT and LDE, as in eng_batch_from_values. */
eng_status eng_lde_peer_host(const uint64_t *const *cols_host, uint32_t num_polys, uint32_t log_n, uint32_t rate_bits,
                             int32_t is_values, uint32_t log_row_shards, uint64_t *coeffs_out_dev, uint64_t *scratch_dev,
                             uint64_t *const *shard_out, uint32_t first_shard);
/* Exchange buffers: device memory outside the stream-ordered pool, exportable to the other ranks of the box through a
 * 64-byte CUDA IPC handle (cudaIpcGetMemHandle / cudaIpcOpenMemHandle). */
eng_status eng_peer_buffer_alloc(uint64_t num_elems, uint64_t **dev_out, uint8_t handle_out[64]);
eng_status eng_peer_buffer_open(const uint8_t handle[64], uint64_t **dev_out);
eng_status eng_peer_buffer_close(uint64_t *peer_ptr);
eng_status eng_peer_buffer_free(uint64_t *dev_ptr);

/* ---- a3: MerkleTree::new(leaves, cap_height)  [plonky2:hash/merkle_tree.rs] ----
 * leaves_host is row-major [num_leaves][leaf_len].  ENG_ERR_INVALID when cap_height > log2(num_leaves) or
 * num_leaves is not a power of two (plonky2 panics). */
eng_status eng_merkle_new(const uint64_t *leaves_host, uint64_t num_leaves, uint32_t leaf_len, uint32_t cap_height,
                          eng_batch **out);
/* Leaves already on the device: element (row, col) at data_dev[row*row_stride + col*col_stride]; not copied, the
 * caller keeps data_dev alive for the life of the handle. */
eng_status eng_merkle_new_dev(const uint64_t *data_dev, uint64_t row_stride, uint64_t col_stride, uint64_t num_leaves,
                              uint32_t leaf_len, uint32_t cap_height, eng_batch **out);

/* ---- accessors (PolynomialBatch fields, MerkleTree::{cap, get, prove}, get_lde_values) ---- */
typedef struct {
    uint32_t num_polys;   /* polynomials.len() (0 for a bare MerkleTree) */
    uint32_t degree_log;  /* PolynomialBatch::degree_log */
    uint32_t rate_bits;
    uint32_t cap_height;
    uint32_t blinding;
    uint32_t leaf_len;    /* num_polys (+4 salt) or MerkleTree leaf length */
    uint64_t num_leaves;  /* 2^(degree_log + rate_bits) */
    uint64_t num_digests; /* 2 * (num_leaves - 2^cap_height) */
} eng_batch_info_t;
eng_status eng_batch_info(const eng_batch *b, eng_batch_info_t *out);
eng_status eng_batch_cap(const eng_batch *b, uint64_t *out_host);                    /* [2^cap_height][4] */
eng_status eng_batch_digests(const eng_batch *b, uint64_t *out_host);                /* plonky2's `digests` order */
eng_status eng_batch_coeffs(const eng_batch *b, uint32_t poly, uint64_t *out_host);  /* polynomials[poly].coeffs */
/* merkle_tree.leaves[first .. first+count], row-major [count][leaf_len] */
eng_status eng_batch_leaves(const eng_batch *b, uint64_t first, uint64_t count, uint64_t *out_host);
/* get_lde_values(index, step) = leaves[bitrev(index*step)] without the salt: num_polys elements */
eng_status eng_batch_lde_values(const eng_batch *b, uint64_t index, uint64_t step, uint64_t *out_host);
/* merkle_tree.prove(leaf_index).siblings, bottom-up; *num_siblings = log2(num_leaves) - cap_height */
eng_status eng_batch_merkle_path(const eng_batch *b, uint64_t leaf_index, uint64_t *siblings_host, uint32_t *num_siblings);
/* Device views for device-side consumers (quotient, FRI): lde is [leaf_len][num_leaves] column-major in
 * bit-reversed row order; coeffs is [num_polys][2^degree_log]. */
eng_status eng_batch_device_ptrs(const eng_batch *b, const uint64_t **lde_dev, const uint64_t **coeffs_dev,
                                 const uint64_t **digests_dev, const uint64_t **cap_dev);
/* Stage times of the call that built the batch, ms, in the order of plonky2's timed! labels:
 * [0] "IFFT"  [1] "FFT + blinding"  [2] "transpose LDEs" (always 0: the layout makes it implicit)
 * [3] "build Merkle tree" leaf hashing  [4] "build Merkle tree" digest levels + cap  [5] host->device copies */
eng_status eng_batch_stage_ms(const eng_batch *b, float out[6]);

/* ---- a9: Challenger<F, PoseidonHash>  [plonky2:iop/challenger.rs] ----
 * Host-side duplex transcript (one permutation per 8 observed elements), bit-compatible with plonky2's: the state
 * crosses the FFI as 30 words: [0..12) sponge_state, [12] input_buffer.len(), [13..21) input_buffer,
 * [21] output_buffer.len(), [22..30) output_buffer. */
typedef struct eng_challenger eng_challenger;
eng_status eng_challenger_new(eng_challenger **out);
eng_status eng_challenger_free(eng_challenger *c);
eng_status eng_challenger_observe(eng_challenger *c, const uint64_t *elements, size_t n);   /* observe_elements */
eng_status eng_challenger_get_challenges(eng_challenger *c, uint64_t *out, size_t n);       /* get_n_challenges */
eng_status eng_challenger_get_state(const eng_challenger *c, uint64_t out[30]);
eng_status eng_challenger_set_state(eng_challenger *c, const uint64_t in[30]);

/* ---- a7: OpeningSet::new's eval_commitment(z, batch)  [plonky2:plonk/proof.rs] ----
 * Evaluates every polynomial of the batch at z = z[0] + z[1]*X in F_p^2; out_host is [num_polys][2]. */
eng_status eng_batch_eval_ext(const eng_batch *b, const uint64_t z[2], uint64_t *out_host);

/* ---- a8: PolynomialBatch::prove_openings -> fri_proof  [plonky2:fri/oracle.rs, fri/prover.rs] ----
 * instance (FriInstanceInfo): [num_batches] then per batch [point.a, point.b, num_polys, (oracle_index << 32 |
 *   polynomial_index) x num_polys].
 * params (FriParams): [degree_bits, rate_bits, cap_height, proof_of_work_bits, num_query_rounds,
 *   reduction_arity_bits.len(), reduction_arity_bits...]  (arity 16 only: ConstantArityBits(4, _)).
 * The challenger must be in the state right after observing the openings; it is advanced exactly as plonky2's.
 * Output (FriProof), a malloc'ed flat u64 array released with eng_blob_free (plonky2's byte format: eng_proof_to_bytes):
 *   [R] R x { [len] cap } | [F] F x (a, b) final_poly | pow_witness |
 *   [Q] Q x { [O] O x { [leaf_len] leaf | [path_len] path x 4 } | [R] R x { [arity] evals x 2 | [path_len] path x 4 } }
 * pow_witness is the SMALLEST valid witness (plonky2's rayon find_any returns an arbitrary one). */
eng_status eng_fri_prove_openings(const uint64_t *instance, const eng_batch *const *oracles, uint32_t num_oracles,
                                  eng_challenger *challenger, const int32_t *params, uint64_t **blob_out, size_t *blob_len);
eng_status eng_blob_free(uint64_t *blob);

/* ---- circuit handle for the plonk rows ----
 * blob: what the rows need of CommonCircuitData / ProverOnlyCircuitData.  Two layouts are accepted.
 * Version 2 (gates are DATA -- any gate set, see csrc/gate_vm.h and INTEGRATION.md):
 *   [0] 0x32424B4C50 ("PLKB2")  [1] total length in words
 *   [2..14) degree_bits, num_wires, num_routed_wires, num_constants (gate constants, without selectors), num_selectors,
 *           num_challenges, quotient_degree_factor, rate_bits, cap_height, proof_of_work_bits, num_query_rounds, num_gates
 *   [14] num_imm  [15] prog_words  [16..20) circuit_digest
 *   num_gates x 12 words (gates in plonky2's degree-sorted order; the index of a gate is the selector value enabling it):
 *     kind, selector_index, group.start, group.end, p0, p1, p2, p3, num_constraints, prog_offset, prog_len, 0
 *   prog_words instruction words (the gates' eval_unfiltered as straight-line programs: add / sub / mul / mad / msub / mov /
 *     emit over wires, gate constants, immediates and 64 registers)   |   num_imm immediates (the first 4 are reserved for
 *     public_inputs_hash).
 *   kind names the gate for the library (prog_len = 0: the program is built by csrc/gate_lib.h): 0 Noop, 1 Constant{p0},
 *   2 PublicInput, 3 Arithmetic{p0 ops}, 4 Poseidon, 5 BaseSum<p0>{p1 limbs}, 6 ArithmeticExtension{p0}, 7 MulExtension{p0},
 *   8 Reducing{p0 coeffs}, 9 ReducingExtension{p0}, 10 RandomAccess{p0 bits, p1 copies, p2 extra constants},
 *   11 Exponentiation{p0 bits}, 12 PoseidonMds, 13 U32Arithmetic{p0}, 14 U32AddMany{p0 addends, p1 ops}, 15 U32Subtraction{p0},
 *   16 U32RangeCheck{p0 limbs}, 17 Comparison{p0 bits, p1 chunks}, 18 CosetInterpolation{p0 subgroup bits, p1 degree},
 *   19 U32Interleave{p0 ops}, 20 UninterleaveToU32{p0 ops}, 21 UninterleaveToB32{p0 ops}; 255 = custom (the program is mandatory).  A library gate whose program is the library's own runs through the evaluators
 *   compiled from the same source (option quot_native_gates); anything else through the bytecode interpreter.
 * Version 1 (the five core gates, kept for round-1 callers): the 12 header words, (kind, selector_index, group.start,
 *   group.end) x num_gates, circuit_digest x 4.
 * constants_sigmas: the batch committed by build() (selectors, gate constants, sigmas -- in that column order);
 * sigma_cols_host: the sigma polynomials' values on the subgroup (prover_data.sigmas, column j = routed wire j). */
typedef struct eng_circuit eng_circuit;
eng_status eng_circuit_new(const uint64_t *blob, const eng_batch *constants_sigmas, const uint64_t *const *sigma_cols_host,
                           eng_circuit **out);
eng_status eng_circuit_free(eng_circuit *c);
/* Row f2, host half of build(): the sigma polynomials' values from the copy constraints (plonky2's Forest +
 * WirePartition::get_sigma_polys).  copies: num_copies x (row_a, column_a, row_b, column_b) over routed wires; the wires
 * of an equivalence class map cyclically onto each other in (row, column) order; sigmas_out [num_routed][2^degree_bits]. */
eng_status eng_build_sigmas(uint32_t degree_bits, uint32_t num_routed, const uint32_t *copies, size_t num_copies, uint64_t *sigmas_out);
typedef struct {
    uint32_t degree_bits, num_wires, num_routed_wires, num_constants /* selectors + gate constants */, num_selectors,
             num_challenges, quotient_degree_factor, num_partial_products, rate_bits, cap_height, num_gates,
             num_gate_constraints;
} eng_circuit_info_t;
eng_status eng_circuit_info(const eng_circuit *c, eng_circuit_info_t *out);
/* Version-2 description of a circuit whose gates the library knows.  header12: the 12 header words above; gates8: num_gates x
 * (kind, selector_index, group.start, group.end, p0, p1, p2, p3); digest4: circuit_digest.  *blob_out: malloc'ed
 * (eng_blob_free).  Host code, no device needed. */
eng_status eng_circuit_describe(const uint64_t *header12, const uint64_t *gates8, uint32_t num_gates, const uint64_t *digest4,
                                uint64_t **blob_out, size_t *blob_len);
/* ---- f2: prover-data cache (CircuitBuilder::build()'s constants||sigmas commitment, not re-paid per process) ----
 * eng_circuit_save writes the circuit description, the constant / sigma polynomials and the cap of their commitment to
 * `path`; eng_circuit_load rebuilds the device-resident commitment from them (LDE + Merkle tree on the GPU, checked against
 * the stored cap) and returns a circuit handle that owns it. */
eng_status eng_circuit_save(const eng_circuit *c, const uint64_t *blob, const char *path);
eng_status eng_circuit_load(const char *path, eng_circuit **out, uint64_t **blob_out, size_t *blob_len);
/* the constants||sigmas batch a circuit handle refers to (owned by the handle after eng_circuit_load) */
eng_status eng_circuit_constants_sigmas(const eng_circuit *c, const eng_batch **out);

/* ---- a5: all_wires_permutation_partial_products  [plonky2:plonk/prover.rs] ----
 * wire_cols_host: the full witness (num_wires columns, only the routed ones are read).  out_host:
 * [num_challenges * (1 + num_partial_products)][n] in the order prove() commits them: Z_0.., pp_0[..], pp_1[..]. */
eng_status eng_partial_products(const eng_circuit *c, const uint64_t *const *wire_cols_host, const uint64_t *betas,
                                const uint64_t *gammas, uint64_t *out_host);

/* ---- a6: compute_quotient_polys + "split up" + "commit to quotient polys"  [plonky2:plonk/prover.rs] ----
 * Returns the committed batch of num_challenges * quotient_degree_factor chunk polynomials. */
eng_status eng_quotient(const eng_circuit *c, const eng_batch *wires, const eng_batch *zs_partial_products,
                        const uint64_t *public_inputs_hash, const uint64_t *betas, const uint64_t *gammas,
                        const uint64_t *alphas, eng_batch **quotient_out);

/* ---- prove_with_partition_witness after witness generation  [plonky2:plonk/prover.rs] ----
 * Everything between "compute full witness" and the returned proof, device-resident between the stages.  blob:
 * wires_cap | zs_partial_products_cap | quotient_cap (2^h x 4 each) | OpeningSet as (a, b) pairs: constants,
 * plonk_sigmas, wires, plonk_zs, plonk_zs_next, partial_products, quotient_polys | FriProof (eng_fri_prove_openings).
 * stage_ms (may be NULL, 8 floats): wires commitment, partial products, Z commit, quotient (values + iNTT), quotient
 * commit, opening set, opening proofs (FRI), total. */
eng_status eng_prove(const eng_circuit *c, const uint64_t *const *wire_cols_host, const uint64_t *public_inputs_hash,
                     uint64_t **blob_out, size_t *blob_len, float *stage_ms);

/* ---- multi-GPU prover: prove_with_partition_witness over the GPUs of one box, one process per GPU ----
 * Every committed batch (constants||sigmas, wires, Z||partial products, quotient chunks) is column-sharded for the
 * iNTT / LDE and row-sharded for hashing (eng_lde_peer_dev / eng_merkle_new_dev); after the column -> row exchange rank g
 * holds the leaf matrix [cols][L/G] of each batch.  The steps that read whole ROWS -- gate constraints + permutation checks
 * of the quotient, the FRI combination -- run on those leaf matrices without further exchange:
 *   x -> w_n x moves an LDE position inside its row shard as long as G <= 2^quotient_degree_bits, so Z(w_n x) is local;
 *   16 consecutive positions (one FRI coset / Merkle leaf of the first commit-phase layer) never straddle row shards.
 * Host orchestration (transcript, all-gathers of the quotient values / layer 0 / openings): eth-lc-plonky2_b200/parallel.py
 * ShardedProver; the proof it assembles is bit-identical to eng_prove's.  Replaces the rayon parallelism inside
 * plonky2::plonk::prover::prove_with_partition_witness (/root/reference/eth-lc-plonky2/src/main.rs:230). */
eng_status eng_circuit_new_sharded(const uint64_t *blob, const uint64_t *const *sigma_cols_host, eng_circuit **out);
/* a5 with the result left on the device: out_dev [num_challenges * (1 + num_partial_products)][n] */
eng_status eng_partial_products_dev(const eng_circuit *c, const uint64_t *const *wire_cols_host, const uint64_t *betas,
                                    const uint64_t *gammas, uint64_t *out_dev);
/* the same from routed wire values already on the device (the sharded prover all-gathers them from the ranks' column shards) */
eng_status eng_partial_products_from_dev(const eng_circuit *c, const uint64_t *wires_dev, const uint64_t *betas, const uint64_t *gammas,
                                         uint64_t *out_dev);
/* host columns -> one device array [num_cols][n] through the engine's pinned staging pipeline */
eng_status eng_h2d_columns(const uint64_t *const *cols_host, uint32_t num_cols, uint64_t n, uint64_t *dst_dev);
/* a6 on row shard `shard` of 2^log_shards: quotient values at LDE positions [shard * L/G, (shard + 1) * L/G) from this rank's
 * leaf matrices ([cols][L/G] column-major) -> out_dev [num_challenges][L/G], position order.  Needs quotient_degree_bits ==
 * rate_bits and G <= 2^quotient_degree_bits. */
eng_status eng_quotient_values_shard_dev(const eng_circuit *c, const uint64_t *cs_rows_dev, const uint64_t *wires_rows_dev,
                                         const uint64_t *zs_rows_dev, uint32_t log_shards, uint32_t shard, const uint64_t *public_inputs_hash,
                                         const uint64_t *betas, const uint64_t *gammas, const uint64_t *alphas, uint64_t *out_dev);
/* gathered shards [G][num_challenges][L/G] -> natural order -> coset iNTT -> chunk coefficients [num_challenges * qdf][n] */
eng_status eng_quotient_coeffs_from_shards_dev(const eng_circuit *c, const uint64_t *gathered_dev, uint32_t log_shards, uint64_t *coeffs_out_dev);
/* a7 on a device coefficient array [num_polys][2^log_n] (a rank's column shard) -> out_host [num_polys][2] */
eng_status eng_eval_ext_dev(const uint64_t *coeffs_dev, uint32_t num_polys, uint32_t log_n, const uint64_t z[2], uint64_t *out_host);
/* a8, first half, on a row shard: FRI layer 0 (the combined, divided codeword) at positions [shard * L/G, ...).  instance:
 * as for eng_fri_prove_openings; rows_dev[o] / widths[o]: leaf matrix and width of oracle o; openings_host: the claimed
 * values (a, b) of every polynomial of the instance in instance order; alpha: drawn by the caller.  out_dev [L/G][2]. */
eng_status eng_fri_combine_shard_dev(const uint64_t *instance, const uint64_t *const *rows_dev, const uint32_t *widths, uint32_t num_oracles,
                                     const uint64_t *openings_host, const uint64_t alpha[2], uint32_t log_l, uint32_t log_shards,
                                     uint32_t shard, uint64_t *out_dev);
/* a8, second half, from the gathered layer 0 ([L][2], position order): commit phase, final polynomial, proof of work, query
 * indices, commit-phase openings.  Blob = FriProof layout with 0 initial oracles per query round; the caller opens rows
 * query_indices_out[0 .. num_query_rounds) of the initial oracles on the ranks that own them and splices them in. */
eng_status eng_fri_prove_from_layer_dev(const uint64_t *layer0_dev, eng_challenger *ch, const int32_t *params, uint64_t **blob_out,
                                        size_t *blob_len, uint64_t *query_indices_out);

/* ---- f4: CircuitData::verify and ProofWithPublicInputs::{to_bytes, from_bytes}  [plonky2:plonk/verifier.rs,
 * util/serialization]; reached from /root/reference/eth-lc-plonky2/src/main.rs:233 ----
 * Host code, as in the reference (verification is a few hundred permutations): no device and no eng_init needed.
 * circuit_blob: the circuit description (either version); constants_sigmas_cap: verifier_only.constants_sigmas_cap
 * (2^cap_height x 4); proof_blob: as returned by eng_prove.  ENG_OK = the proof verifies; ENG_ERR_INVALID = rejected, with
 * the failing check in eng_last_error (the Rust shim turns that into verify()'s Err). */
eng_status eng_verify(const uint64_t *circuit_blob, const uint64_t *constants_sigmas_cap, const uint64_t *public_inputs_hash,
                      const uint64_t *proof_blob, size_t proof_len);
/* public_inputs_hash = PoseidonHash::hash_no_pad(public_inputs) (host code): the 4 words eng_prove / eng_verify take. */
eng_status eng_public_inputs_hash(const uint64_t *public_inputs, size_t num_public_inputs, uint64_t out[4]);
/* plonky2's wire format as restated (to be re-checked against the source): little-endian
 * canonical u64 per field element, caps / openings / FRI layers in struct order without length prefixes, one byte of
 * sibling count in front of every Merkle proof, the public inputs last.  *bytes_out: malloc'ed (eng_bytes_free). */
eng_status eng_proof_to_bytes(const uint64_t *circuit_blob, const uint64_t *proof_blob, size_t proof_len, const uint64_t *public_inputs,
                              size_t num_public_inputs, uint8_t **bytes_out, size_t *bytes_len);
eng_status eng_proof_from_bytes(const uint64_t *circuit_blob, const uint8_t *bytes, size_t bytes_len, uint64_t **proof_blob_out,
                                size_t *proof_len, uint64_t **public_inputs_out, size_t *num_public_inputs);
eng_status eng_bytes_free(uint8_t *bytes);

/* Synthetic circuits (tests / bench input generators, host code): 2^degree_bits rows of plonky2 gates with a satisfying
 * witness and non-trivial copy constraints.  Caller-allocated outputs, column-major.
 * Version 1: the five core gates; constants [4][n] (selector 0, selector 1, gate constants 0 and 1), sigmas [80][n], wires
 * [135][n], pi_hash [4], circuit_blob [36] (version-1 layout). */
eng_status eng_synth_circuit(uint32_t degree_bits, uint64_t seed, uint64_t *constants, uint64_t *sigmas, uint64_t *wires,
                             uint64_t *pi_hash, uint64_t *circuit_blob);
/* Version 2: the gates selected by kinds_mask (bit k = gate kind k of the list above; Noop and PublicInput always), sorted by
 * degree and grouped into selector polynomials by plonky2's greedy rule.  constants: [ENG_SYNTH_MAX_CONSTANTS][n] of which
 * the first *num_constants columns are used (selectors, then the two gate constants); *blob_out: malloc'ed version-2
 * description (eng_blob_free). */
#define ENG_SYNTH_MAX_CONSTANTS 8
eng_status eng_synth_circuit_v2(uint32_t degree_bits, uint64_t seed, uint64_t kinds_mask, uint64_t *constants, uint64_t *sigmas,
                                uint64_t *wires, uint64_t *pi_hash, uint32_t *num_constants, uint64_t **blob_out, size_t *blob_len);

#ifdef __cplusplus
}
#endif
#endif
