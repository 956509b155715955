// Poseidon Merkle tree: leaf hashing and the 2-to-1 digest tree down to the cap.
//
// Replaces plonky2::hash::merkle_tree::MerkleTree::new (fill_digests_buf / fill_subtree) and
// PoseidonHash::{hash_or_noop, two_to_one} (dep plonky2 0.1.4, /root/reference/Cargo.lock:2347-2350;
// SURVEY.md 3.3, A.5, A.6).  The `digests` vector keeps plonky2's interleaved layout bit for bit
// (a node is stored next to its sibling inside its parent's region; roots only in `cap`), so
// MerkleTree::prove is the same index formula and the export of `digests` is a plain copy.
//
// B200 design: one sponge state per thread.  Leaves are addressed through (row_stride, col_stride), so the
// same kernel hashes the engine's column-major LDE (coalesced: consecutive threads read consecutive rows of one
// column) and row-major leaves handed in through MerkleTree::new.
#pragma once
#include "poseidon.cuh"

struct MerkleParams {
    const u64 *data;      // leaf elements: element (row, col) at data[row*row_stride + col*col_stride]
    u64 row_stride, col_stride;
    u32 width;            // elements per leaf
    u32 noop_max;         // leaves of at most this many elements are copied, not hashed (hash_or_noop: 4)
    u64 num_leaves;       // power of two
    u32 num_layers;       // log2(num_leaves) - cap_height
    u64 *digests;         // [2*(num_leaves - 2^cap_height)][4]
    u64 *cap;             // [2^cap_height][4]
};

// Position (in digests, units of one digest) of node `g` (global index across all cap subtrees) of `layer`
// (0 = leaf digests).  Pair p of layer i inside a subtree sits at 2*((p << (i+1)) + 2^i - 1)  [merkle_tree.rs::prove].
GL_HD u64 merkle_digest_pos(u32 num_layers, u32 layer, u64 g) {
    u32 per_tree_log = num_layers - layer;
    u64 t = g >> per_tree_log, jj = g & (((u64)1 << per_tree_log) - 1);
    u64 sub = ((u64)2 << num_layers) - 2;
    return t * sub + 2 * (((jj >> 1) << (layer + 1)) + ((u64)1 << layer) - 1) + (jj & 1);
}

GL_HD void merkle_store_digest(const MerkleParams &p, u32 layer, u64 g, const u64 d[4]) {
    u64 *dst = (layer == p.num_layers) ? p.cap + 4 * g : p.digests + 4 * merkle_digest_pos(p.num_layers, layer, g);
#ifdef __CUDA_ARCH__
    reinterpret_cast<ulonglong2 *>(dst)[0] = make_ulonglong2(d[0], d[1]);
    reinterpret_cast<ulonglong2 *>(dst)[1] = make_ulonglong2(d[2], d[3]);
#else
    dst[0] = d[0]; dst[1] = d[1]; dst[2] = d[2]; dst[3] = d[3];
#endif
}

// hash_or_noop(leaf j): <= 4 elements are copied (zero padded), otherwise overwrite-mode sponge at rate 8.
GL_HD void merkle_hash_leaf(const MerkleParams &p, u64 j) {
    const u64 *row = p.data + j * p.row_stride;
    u64 d[4];
    if (p.width <= p.noop_max) {
#pragma unroll
        for (u32 i = 0; i < 4; i++) d[i] = i < p.width ? gl_canon(row[i * p.col_stride]) : 0;
    } else {
        u64 s[12];
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = 0;
        // one (inlined) copy of the permutation: the tail chunk overwrites fewer lanes
        u32 chunks = (p.width + 7) / 8;
        PSD_UNROLL1
        for (u32 c = 0; c < chunks; c++) {
            const u64 *src = row + (u64)(8 * c) * p.col_stride;
            u32 len = p.width - 8 * c;
#pragma unroll
            for (int i = 0; i < 8; i++)
                if ((u32)i < len) s[i] = src[i * p.col_stride];
            poseidon_permute(s);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) d[i] = gl_canon(s[i]);
    }
    merkle_store_digest(p, 0, j, d);
}

// parent g of layer+1 = two_to_one(children 2g, 2g+1 of `layer`)
GL_HD void merkle_hash_node(const MerkleParams &p, u32 layer, u64 g) {
    const u64 *ch = p.digests + 4 * merkle_digest_pos(p.num_layers, layer, 2 * g);
    u64 d[4];
    poseidon_two_to_one(ch, ch + 4, d);
    merkle_store_digest(p, layer + 1, g, d);
}

#ifdef __CUDACC__
__global__ void __launch_bounds__(128, 4) merkle_leaves_kernel(MerkleParams p) {
    u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < p.num_leaves) merkle_hash_leaf(p, j);
}
__global__ void __launch_bounds__(128, 4) merkle_level_kernel(MerkleParams p, u32 layer) {
    u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < (p.num_leaves >> (layer + 1))) merkle_hash_node(p, layer, g);
}
#endif
