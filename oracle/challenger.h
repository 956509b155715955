/* ORACLE -- TEST INFRASTRUCTURE ONLY (see gl.h header).  PARITY UNPINNED (restated; source absent).
 * Challenger: Poseidon duplex transcript [DEP plonky2:iop/challenger.rs]. */
#ifndef ORACLE_CHALLENGER_H
#define ORACLE_CHALLENGER_H
#include "poseidon.h"

struct OrcChallenger {
    u64 state[12];
    u64 in_buf[8];  int n_in;
    u64 out_buf[8]; int n_out;
};
static inline void orc_ch_init(OrcChallenger *c) { memset(c, 0, sizeof(*c)); }
/* duplexing: overwrite state[0..len] with the buffered inputs, permute, outputs = state[0..8] */
static inline void orc_ch_duplex(OrcChallenger *c) {
    for (int i = 0; i < c->n_in; i++) c->state[i] = c->in_buf[i];
    c->n_in = 0;
    orc_poseidon(c->state);
    for (int i = 0; i < 8; i++) c->out_buf[i] = c->state[i];
    c->n_out = 8;
}
static inline void orc_ch_observe(OrcChallenger *c, u64 e) {
    c->n_out = 0;
    c->in_buf[c->n_in++] = gl_canon(e);
    if (c->n_in == 8) orc_ch_duplex(c);
}
static inline void orc_ch_observe_n(OrcChallenger *c, const u64 *e, size_t n) { for (size_t i = 0; i < n; i++) orc_ch_observe(c, e[i]); }
static inline void orc_ch_observe_ext(OrcChallenger *c, gl2 e) { orc_ch_observe(c, e.a); orc_ch_observe(c, e.b); }
/* get_challenge: pops from the END of the output buffer */
static inline u64 orc_ch_challenge(OrcChallenger *c) {
    if (c->n_in > 0 || c->n_out == 0) orc_ch_duplex(c);
    return c->out_buf[--c->n_out];
}
static inline gl2 orc_ch_challenge_ext(OrcChallenger *c) { u64 a = orc_ch_challenge(c); u64 b = orc_ch_challenge(c); return gl2_make(a, b); }
#endif
