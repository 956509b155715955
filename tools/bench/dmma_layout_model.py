# Warp-level model of the DMMA Poseidon layout of tools/bench/poseidon_dmma.cuh (development tool): field arithmetic with exact
# integers, the mma.m8n8k4 fragment maps, the rotated four-batch lane map, the B fragments shared by all batches, T0 through
# the spare column and the merged rank-one term.  Prints "ok" when 32 states equal the reference permutation.
import os, sys, random
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests", "golden"))
from plonk_restatement import P, RC, CIRC, poseidon

def circ2():
    return [sum(CIRC[i] * CIRC[(m - i) % 12] for i in range(12)) for m in range(12)]
C2 = circ2()

def dmma(A, B, C0, C1):
    D0 = [0] * 32; D1 = [0] * 32
    for g in range(8):
        for t in range(4):
            for e, (Cc, D) in enumerate(((C0, D0), (C1, D1))):
                n = 2 * t + e
                acc = Cc[4 * g + t]
                for k in range(4):
                    acc += A[4 * g + k] * B[4 * n + k]
                D[4 * g + t] = acc
    return D0, D1

def lane_of(b, c, t): return (4 * c + t + b) % 12
R_T = [0, 9, 10, 11]          # virtual row delivered to thread t' through the odd columns of tile 1
BSTAR = [0, 3, 2, 1]          # batch whose lane 0 thread t holds
CSTAR = [0, 2, 2, 2]          # ... and its slot

def bfrag(cc, c, tile, extra):
    # thread (g, t) holds B[k = t][n = g]
    out = [0] * 32
    for g in range(8):
        for t in range(4):
            tp, e = g >> 1, g & 1
            if tile == 0:
                v_out = 4 * e + tp
                val = cc[(4 * c + t - v_out) % 12]
            else:
                if e == 0:
                    val = cc[(4 * c + t - (8 + tp)) % 12]
                else:
                    val = CIRC[(4 * c + t - R_T[tp]) % 12] if extra else 0
            out[4 * g + t] = val
    return out

def layer(S, cc, init_fn, extra, t0c=None):
    # S[b][c][lane]; returns new S and the extra column E[b][lane]
    N = [[None] * 3 for _ in range(4)]
    E = [None] * 4
    for b in range(4):
        c0 = [init_fn(lane_of(b, 0, l & 3)) for l in range(32)]
        c1 = [init_fn(lane_of(b, 1, l & 3)) for l in range(32)]
        c2 = [init_fn(lane_of(b, 2, l & 3)) for l in range(32)]
        cx = [t0c if t0c is not None else 0] * 32
        d00, d01 = c0, c1
        d10, d11 = c2, cx
        for c in range(3):
            d00, d01 = dmma(S[b][c], bfrag(cc, c, 0, extra), d00, d01)
            d10, d11 = dmma(S[b][c], bfrag(cc, c, 1, extra), d10, d11)
        N[b][0], N[b][1], N[b][2], E[b] = d00, d01, d10, d11
    return N, E

def permute_warp(states):
    # states[b][g] = list of 12 field elements
    S = [[[states[b][l >> 2][lane_of(b, c, l & 3)] for l in range(32)] for c in range(3)] for b in range(4)]
    # + RC_0
    for b in range(4):
        for c in range(3):
            for l in range(32):
                S[b][c][l] = (S[b][c][l] + RC[lane_of(b, c, l & 3)]) % P
    def full(S, L):
        nxt = L + 1 if L < 4 else 27 + (L - 4)
        X = [[[pow(v, 7, P) for v in S[b][c]] for c in range(3)] for b in range(4)]
        N, _ = layer(X, CIRC, lambda ln: RC[12 * nxt + ln] if nxt < 30 else 0, False)
        for b in range(4):
            for c in range(3):
                for l in range(32):
                    if lane_of(b, c, l & 3) == 0:
                        N[b][c][l] += 8 * X[b][c][l]
                    N[b][c][l] %= P
        return N
    for L in range(4):
        S = full(S, L)
    M = [[0] * 12 for _ in range(12)]
    for r in range(12):
        for i in range(12):
            M[r][(i + r) % 12] += CIRC[i]
    M[0][0] += 8
    for p in range(11):
        r1, r2 = 4 + 2 * p + 1, 4 + 2 * p + 2
        t0c = RC[12 * r1]
        K = [(RC[12 * r2 + r] + sum(M[r][j] * RC[12 * r1 + j] for j in range(12))) % P for r in range(12)]
        K[0] = (K[0] - 8 * t0c) % P
        # each thread: lane 0 of its batch
        W0 = [0] * 32
        for l in range(32):
            t = l & 3
            a = S[BSTAR[t]][CSTAR[t]][l]
            W0[l] = pow(a, 7, P)
            S[BSTAR[t]][CSTAR[t]][l] = W0[l]
        N, E = layer(S, C2, lambda ln: K[ln], True, t0c)
        U = [0] * 32; B7 = [0] * 32
        for l in range(32):
            t = l & 3
            T0 = (E[BSTAR[t]][l] + 8 * W0[l]) % P
            bp = pow(T0, 7, P)
            U[l] = (8 * W0[l] + bp - T0) % P
            B7[l] = bp
        for b in range(4):
            tb = (4 - b) % 4
            for c in range(3):
                for l in range(32):
                    u = U[(l & ~3) | tb]         # shfl
                    ln = lane_of(b, c, l & 3)
                    v = N[b][c][l] + u * CIRC[(-ln) % 12]
                    if ln == 0:
                        v += 8 * B7[l]
                    N[b][c][l] = v % P
        S = N
    for L in range(4, 8):
        S = full(S, L)
    out = [[[0] * 12 for g in range(8)] for b in range(4)]
    for b in range(4):
        for c in range(3):
            for l in range(32):
                out[b][l >> 2][lane_of(b, c, l & 3)] = S[b][c][l]
    return out

random.seed(1)
states = [[[random.randrange(P) for _ in range(12)] for g in range(8)] for b in range(4)]
out = permute_warp(states)
ok = all(out[b][g] == poseidon(states[b][g]) for b in range(4) for g in range(8))
print("ok" if ok else "MISMATCH")
