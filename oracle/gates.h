/* ORACLE -- TEST INFRASTRUCTURE ONLY (see gl.h header).  PARITY UNPINNED (tier C of SURVEY.md Appendix A).
 *
 * Direct restatement of Gate::eval_unfiltered for the gates beyond the five core ones: what the recursion verifier
 * (/root/reference/eth-lc-plonky2/src/targets.rs:468-470), plonky2_crypto's SHA-256 / biguint gadgets
 * (/root/reference/eth-lc-plonky2/src/merkle_tree_gadget.rs:37, targets.rs:198,317) and split_le_base::<2>
 * (/root/reference/eth-lc-plonky2/src/utils.rs:102-103) put into the eth-lc circuit.
 *   [DEP plonky2:gates/{base_sum,arithmetic_extension,multiplication_extension,reducing,reducing_extension,random_access,
 *        exponentiation,poseidon_mds,coset_interpolation}.rs]
 *   [DEP plonky2_crypto @3f71378 (plonky2_u32 gates):gates/{arithmetic_u32,add_many_u32,subtraction_u32,range_check_u32,
 *        comparison,interleave_u32,uninterleave_to_u32,uninterleave_to_b32}.rs]
 * Written as formulas over a generic field (O = OpsBase for the prover's points, OpsExt for the verifier's zeta); the engine
 * evaluates the same gates from BYTECODE (eth-lc-plonky2_b200/csrc/gate_lib.h), so agreement of the two is a cross-check of
 * two independently written forms.  Wire layouts are spelled out per gate below (not shared with the product).
 */
#ifndef ORACLE_GATES_H
#define ORACLE_GATES_H
#include <vector>
#include "gl.h"
#include "poseidon.h"

enum {
    ORC_G_BASE_SUM = 5, ORC_G_ARITHMETIC_EXT = 6, ORC_G_MUL_EXT = 7, ORC_G_REDUCING = 8, ORC_G_REDUCING_EXT = 9,
    ORC_G_RANDOM_ACCESS = 10, ORC_G_EXPONENTIATION = 11, ORC_G_POSEIDON_MDS = 12, ORC_G_U32_ARITHMETIC = 13,
    ORC_G_U32_ADD_MANY = 14, ORC_G_U32_SUBTRACTION = 15, ORC_G_U32_RANGE_CHECK = 16, ORC_G_COMPARISON = 17,
    ORC_G_COSET_INTERPOLATION = 18, ORC_G_U32_INTERLEAVE = 19, ORC_G_UNINTERLEAVE_TO_U32 = 20, ORC_G_UNINTERLEAVE_TO_B32 = 21
};

/* an element of the extension algebra spread over two wires */
template <class O> struct OrcPair { typename O::T x, y; };
template <class O> OrcPair<O> orc_pair_at(const typename O::T *w, int first) { OrcPair<O> r = {w[first], w[first + 1]}; return r; }
template <class O> OrcPair<O> orc_pair_mul(OrcPair<O> a, OrcPair<O> b) {   /* (x + yX)(x' + y'X), X^2 = 7 */
    OrcPair<O> r;
    r.x = O::add(O::mul(a.x, b.x), O::scale(O::mul(a.y, b.y), GL_W));
    r.y = O::add(O::mul(a.x, b.y), O::mul(a.y, b.x));
    return r;
}
template <class O> typename O::T orc_limb_range(typename O::T limb, int bound) {   /* limb (limb - 1) ... (limb - bound + 1) */
    typename O::T acc = limb;
    for (int k = 1; k < bound; k++) acc = O::mul(acc, O::sub(limb, O::from((u64)k)));
    return acc;
}
template <class O> typename O::T orc_horner(const typename O::T *v, int stride, int count, u64 base) {   /* sum v[i] base^i */
    typename O::T acc = O::from(0);
    for (int i = count - 1; i >= 0; i--) acc = O::add(O::scale(acc, base), v[i * stride]);
    return acc;
}

/* constraints of the non-core gates, in order.  p = the gate's parameters (see include/plonky2_b200.h). */
template <class O, class Emit>
bool orc_eval_gate_ext(int kind, const int p[4], const typename O::T *w, const typename O::T *c, Emit emit) {
    typedef typename O::T T;
    const T one = O::from(1);
    switch (kind) {
    case ORC_G_BASE_SUM: {   /* wire 0 = sum, wires 1..=num_limbs = limbs (little-endian) */
        const int B = p[0], nl = p[1];
        emit(O::sub(orc_horner<O>(w + 1, 1, nl, (u64)B), w[0]));
        for (int j = 0; j < nl; j++) emit(orc_limb_range<O>(w[1 + j], B));
        return true;
    }
    case ORC_G_ARITHMETIC_EXT: {   /* op i: multiplicands 8i, 8i+2; addend 8i+4; output 8i+6 */
        for (int i = 0; i < p[0]; i++) {
            OrcPair<O> m = orc_pair_mul<O>(orc_pair_at<O>(w, 8 * i), orc_pair_at<O>(w, 8 * i + 2));
            OrcPair<O> ad = orc_pair_at<O>(w, 8 * i + 4), out = orc_pair_at<O>(w, 8 * i + 6);
            emit(O::sub(out.x, O::add(O::mul(m.x, c[0]), O::mul(ad.x, c[1]))));
            emit(O::sub(out.y, O::add(O::mul(m.y, c[0]), O::mul(ad.y, c[1]))));
        }
        return true;
    }
    case ORC_G_MUL_EXT: {   /* op i: multiplicands 6i, 6i+2; output 6i+4 */
        for (int i = 0; i < p[0]; i++) {
            OrcPair<O> m = orc_pair_mul<O>(orc_pair_at<O>(w, 6 * i), orc_pair_at<O>(w, 6 * i + 2)), out = orc_pair_at<O>(w, 6 * i + 4);
            emit(O::sub(out.x, O::mul(m.x, c[0])));
            emit(O::sub(out.y, O::mul(m.y, c[0])));
        }
        return true;
    }
    case ORC_G_REDUCING: case ORC_G_REDUCING_EXT: {
        /* output 0..2, alpha 2..4, old_acc 4..6, coefficients from 6 (1 or 2 wires each), then the intermediate accumulators;
         * the last accumulator IS the output */
        const int nc = p[0], cw = kind == ORC_G_REDUCING ? 1 : 2;
        OrcPair<O> alpha = orc_pair_at<O>(w, 2), acc = orc_pair_at<O>(w, 4);
        for (int i = 0; i < nc; i++) {
            OrcPair<O> t = orc_pair_mul<O>(acc, alpha);
            t.x = O::add(t.x, w[6 + cw * i]);
            if (cw == 2) t.y = O::add(t.y, w[6 + 2 * i + 1]);
            OrcPair<O> nxt = i == nc - 1 ? orc_pair_at<O>(w, 0) : orc_pair_at<O>(w, 6 + cw * nc + 2 * i);
            emit(O::sub(t.x, nxt.x));
            emit(O::sub(t.y, nxt.y));
            acc = nxt;
        }
        return true;
    }
    case ORC_G_RANDOM_ACCESS: {
        /* copy k: access_index (2 + 2^bits) k, claimed element + 1, list + 2 ..; extra constants after the copies; the bits of
         * copy k start at num_routed + k * bits, num_routed = (2 + 2^bits) copies + extra */
        const int bits = p[0], copies = p[1], extra = p[2], vs = 1 << bits, per = 2 + vs;
        const int routed = per * copies + extra;
        for (int k = 0; k < copies; k++) {
            const T *b = w + routed + k * bits;
            for (int i = 0; i < bits; i++) emit(O::mul(b[i], O::sub(b[i], one)));
            emit(O::sub(orc_horner<O>(b, 1, bits, 2), w[per * k]));
            std::vector<T> items(w + per * k + 2, w + per * k + 2 + vs);
            for (int i = 0; i < bits; i++) {
                std::vector<T> nxt;
                for (size_t j = 0; j + 1 < items.size(); j += 2) nxt.push_back(O::add(items[j], O::mul(b[i], O::sub(items[j + 1], items[j]))));
                items.swap(nxt);
            }
            emit(O::sub(items[0], w[per * k + 1]));
        }
        for (int i = 0; i < extra; i++) emit(O::sub(c[i], w[per * copies + i]));
        return true;
    }
    case ORC_G_EXPONENTIATION: {   /* base 0; power bits 1..=n; output n + 1; intermediate values from n + 2 */
        const int n = p[0];
        for (int i = 0; i < n; i++) {
            T prev = i == 0 ? one : O::mul(w[n + 2 + i - 1], w[n + 2 + i - 1]);
            T bit = w[1 + (n - 1 - i)];
            T factor = O::add(O::mul(bit, w[0]), O::sub(one, bit));
            emit(O::sub(O::mul(prev, factor), w[n + 2 + i]));
        }
        emit(O::sub(w[n + 1], w[n + 2 + n - 1]));
        return true;
    }
    case ORC_G_POSEIDON_MDS: {   /* inputs 12 pairs from 0, outputs 12 pairs from 24 */
        for (int r = 0; r < 12; r++)
            for (int h = 0; h < 2; h++) {
                T acc = O::from(0);
                for (int i = 0; i < 12; i++) acc = O::add(acc, O::scale(w[2 * ((i + r) % 12) + h], ORC_MDS_CIRC[i]));
                if (r == 0) acc = O::add(acc, O::scale(w[h], POSEIDON_MDS_DIAG0));
                emit(O::sub(w[24 + 2 * r + h], acc));
            }
        return true;
    }
    case ORC_G_U32_ARITHMETIC: {
        /* op i: multiplicands 6i, 6i+1; addend 6i+2; low, high halves 6i+3, 6i+4; inverse 6i+5; 32 two-bit limbs of the
         * output at 6 ops + 32 i */
        const int ops = p[0];
        for (int i = 0; i < ops; i++) {
            T computed = O::add(O::mul(w[6 * i], w[6 * i + 1]), w[6 * i + 2]);
            T lo = w[6 * i + 3], hi = w[6 * i + 4];
            T hi_not_max = O::sub(O::mul(w[6 * i + 5], O::sub(O::from(0xFFFFFFFFULL), hi)), one);
            emit(O::mul(hi_not_max, lo));
            emit(O::sub(O::add(O::scale(hi, 1ULL << 32), lo), computed));
            const T *limbs = w + 6 * ops + 32 * i;
            for (int j = 31; j >= 0; j--) emit(orc_limb_range<O>(limbs[j], 4));
            emit(O::sub(orc_horner<O>(limbs, 1, 16, 4), lo));
            emit(O::sub(orc_horner<O>(limbs + 16, 1, 16, 4), hi));
        }
        return true;
    }
    case ORC_G_U32_ADD_MANY: {
        /* op i: addends (a + 3) i .., carry in, result, carry out; 16 result limbs + 2 carry limbs at (a + 3) ops + 18 i */
        const int a = p[0], ops = p[1];
        for (int i = 0; i < ops; i++) {
            const T *op = w + (a + 3) * i;
            T computed = op[a];
            for (int j = 0; j < a; j++) computed = O::add(computed, op[j]);
            emit(O::sub(O::add(O::scale(op[a + 2], 1ULL << 32), op[a + 1]), computed));
            const T *limbs = w + (a + 3) * ops + 18 * i;
            for (int j = 17; j >= 0; j--) emit(orc_limb_range<O>(limbs[j], 4));
            emit(O::sub(orc_horner<O>(limbs, 1, 16, 4), op[a + 1]));
            emit(O::sub(orc_horner<O>(limbs + 16, 1, 2, 4), op[a + 2]));
        }
        return true;
    }
    case ORC_G_U32_SUBTRACTION: {   /* op i: x, y, borrow in, result, borrow out at 5 i; 16 limbs at 5 ops + 16 i */
        const int ops = p[0];
        for (int i = 0; i < ops; i++) {
            const T *op = w + 5 * i;
            T initial = O::sub(O::sub(op[0], op[1]), op[2]);
            emit(O::sub(op[3], O::add(initial, O::scale(op[4], 1ULL << 32))));
            const T *limbs = w + 5 * ops + 16 * i;
            for (int j = 15; j >= 0; j--) emit(orc_limb_range<O>(limbs[j], 4));
            emit(O::sub(orc_horner<O>(limbs, 1, 16, 4), op[3]));
            emit(O::mul(op[4], O::sub(one, op[4])));
        }
        return true;
    }
    case ORC_G_U32_RANGE_CHECK: {   /* k input limbs, then 16 two-bit aux limbs for each */
        const int k = p[0];
        for (int i = 0; i < k; i++) {
            const T *aux = w + k + 16 * i;
            emit(O::sub(orc_horner<O>(aux, 1, 16, 4), w[i]));
            for (int j = 0; j < 16; j++) emit(orc_limb_range<O>(aux[j], 4));
        }
        return true;
    }
    case ORC_G_COMPARISON: {
        /* first 0, second 1, result 2, most-significant diff 3, then num_chunks each of: first chunks, second chunks, equality
         * dummies, chunks-equal flags, intermediate values; then chunk_bits + 1 bits of 2^chunk_bits + diff */
        const int nb = p[0], nch = p[1], cb = (nb + nch - 1) / nch;
        const T *fc = w + 4, *sc = w + 4 + nch, *dummy = w + 4 + 2 * nch, *eq = w + 4 + 3 * nch, *inter = w + 4 + 4 * nch, *bits = w + 4 + 5 * nch;
        emit(O::sub(orc_horner<O>(fc, 1, nch, 1ULL << cb), w[0]));
        emit(O::sub(orc_horner<O>(sc, 1, nch, 1ULL << cb), w[1]));
        T msd = O::from(0);
        for (int i = 0; i < nch; i++) {
            emit(orc_limb_range<O>(fc[i], 1 << cb));
            emit(orc_limb_range<O>(sc[i], 1 << cb));
            T diff = O::sub(sc[i], fc[i]);
            emit(O::sub(O::mul(diff, dummy[i]), O::sub(one, eq[i])));
            emit(O::mul(eq[i], diff));
            emit(O::sub(inter[i], O::mul(eq[i], msd)));
            msd = O::add(inter[i], O::mul(O::sub(one, eq[i]), diff));
        }
        emit(O::sub(w[3], msd));
        for (int i = 0; i <= cb; i++) emit(O::mul(bits[i], O::sub(one, bits[i])));
        emit(O::sub(O::add(O::from(1ULL << cb), w[3]), orc_horner<O>(bits, 1, cb + 1, 2)));
        emit(O::sub(w[2], bits[cb]));
        return true;
    }
    case ORC_G_COSET_INTERPOLATION: {
        /* wire 0 = shift; values i at 1 + 2 i (2^bits of them); evaluation point, evaluation value; then the intermediate
         * evals, the intermediate products, and the shifted evaluation point.  Interpolation over the subgroup <g> by the
         * barycentric formula, evaluated chunk by chunk (degree points, then degree - 1 per intermediate pair). */
        const int bits = p[0], deg = p[1], n = 1 << bits, ni = (n - 2) / (deg - 1);
        const int w_point = 1 + 2 * n, w_value = w_point + 2, w_inter = w_value + 2, w_shifted = w_inter + 4 * ni;
        std::vector<u64> xs(n), bw(n);
        const u64 g = gl_root_of_unity(bits);
        xs[0] = 1;
        for (int i = 1; i < n; i++) xs[i] = gl_mul(xs[i - 1], g);
        for (int i = 0; i < n; i++) {
            u64 d = 1;
            for (int j = 0; j < n; j++) if (j != i) d = gl_mul(d, gl_sub(xs[i], xs[j]));
            bw[i] = gl_inv(d);
        }
        const OrcPair<O> z = orc_pair_at<O>(w, w_shifted), pt = orc_pair_at<O>(w, w_point);
        emit(O::sub(pt.x, O::mul(z.x, w[0])));
        emit(O::sub(pt.y, O::mul(z.y, w[0])));
        OrcPair<O> ev = {O::from(0), O::from(0)}, pr = {O::from(1), O::from(0)};
        int next = 0;
        for (int c = 0; c <= ni; c++) {
            if (c > 0) {
                const OrcPair<O> ie = orc_pair_at<O>(w, w_inter + 2 * (c - 1)), ip = orc_pair_at<O>(w, w_inter + 2 * (ni + c - 1));
                emit(O::sub(ie.x, ev.x)); emit(O::sub(ie.y, ev.y));
                emit(O::sub(ip.x, pr.x)); emit(O::sub(ip.y, pr.y));
                ev = ie; pr = ip;
            }
            const int count = c == 0 ? deg : deg - 1;
            for (int k = 0; k < count && next < n; k++, next++) {
                OrcPair<O> term = {O::sub(z.x, O::from(xs[next])), z.y};
                OrcPair<O> v = orc_pair_at<O>(w, 1 + 2 * next);
                OrcPair<O> wv = {O::scale(v.x, bw[next]), O::scale(v.y, bw[next])};
                OrcPair<O> a = orc_pair_mul<O>(ev, term), b = orc_pair_mul<O>(wv, pr);
                ev.x = O::add(a.x, b.x); ev.y = O::add(a.y, b.y);
                pr = orc_pair_mul<O>(pr, term);
            }
        }
        const OrcPair<O> out = orc_pair_at<O>(w, w_value);
        emit(O::sub(out.x, ev.x)); emit(O::sub(out.y, ev.y));
        return true;
    }
    case ORC_G_U32_INTERLEAVE: {   /* op i: x at 2 i, x_interleaved at 2 i + 1; 32 bits at 2 ops + 32 i (little-endian) */
        const int ops = p[0];
        for (int i = 0; i < ops; i++) {
            const T *bits = w + 2 * ops + 32 * i;
            for (int j = 31; j >= 0; j--) emit(O::mul(bits[j], O::sub(one, bits[j])));
            emit(O::sub(w[2 * i], orc_horner<O>(bits, 1, 32, 2)));
            emit(O::sub(w[2 * i + 1], orc_horner<O>(bits, 1, 32, 4)));
        }
        return true;
    }
    case ORC_G_UNINTERLEAVE_TO_U32: case ORC_G_UNINTERLEAVE_TO_B32: {
        /* op i: x_interleaved, evens, odds at 3 i ..; 64 bits at 3 ops + 64 i; evens / odds collect every second bit, packed
         * with base 2 (to U32) or left spread with base 4 (to B32) */
        const int ops = p[0];
        const u64 base = kind == ORC_G_UNINTERLEAVE_TO_U32 ? 2 : 4;
        for (int i = 0; i < ops; i++) {
            const T *bits = w + 3 * ops + 64 * i;
            for (int j = 63; j >= 0; j--) emit(O::mul(bits[j], O::sub(one, bits[j])));
            emit(O::sub(w[3 * i], orc_horner<O>(bits, 1, 64, 2)));
            emit(O::sub(w[3 * i + 1], orc_horner<O>(bits, 2, 32, base)));
            emit(O::sub(w[3 * i + 2], orc_horner<O>(bits + 1, 2, 32, base)));
        }
        return true;
    }
    }
    return false;
}
#endif
