/*
 * vsim_rt.h -- TEST INFRASTRUCTURE ONLY: runtime of the C models that oracle/vsim/v2c.py generates from the
 * reference's Verilog (two-state values, ordered non-blocking update queue for wide vectors).
 */
#ifndef VSIM_RT_H
#define VSIM_RT_H
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define VS_MAX_WIDE 16
typedef struct vs_model vs_model;

/* one queued non-blocking update of a wide vector */
typedef struct {
    int kind;            /* 0 bit, 1 part fill, 2 copy */
    int dst, src;
    int64_t a, b;        /* bit: a = index; part: a = msb, b = lsb */
    uint64_t value;
    size_t snap_off;     /* copy: offset of the source snapshot in snap[] */
} vs_upd;

#define VS_MODEL_HEADER                                                                      \
    int xfill;                                                                               \
    int n_wide;                                                                              \
    uint64_t *wide[VS_MAX_WIDE];                                                             \
    uint64_t wide_w[VS_MAX_WIDE];                                                            \
    vs_upd *q; size_t qn, qcap;                                                              \
    uint64_t *snap; size_t snap_n, snap_cap;

/* the generated struct starts with VS_MODEL_HEADER; this view gives the runtime access to it */
typedef struct { VS_MODEL_HEADER } vs_header;
#define VS_H(m) ((vs_header *)(m))

static inline uint64_t vs_shl(uint64_t a, uint64_t b) { return b >= 64 ? 0ull : a << b; }
static inline uint64_t vs_shr(uint64_t a, uint64_t b) { return b >= 64 ? 0ull : a >> b; }

static inline uint64_t *vs_wide_new(uint64_t nbits, int xfill) {
    const size_t nw = (size_t)((nbits + 63) / 64);
    uint64_t *p = (uint64_t *)malloc((nw ? nw : 1) * 8);
    if (!p) return NULL;
    memset(p, xfill ? 0xFF : 0, (nw ? nw : 1) * 8);
    if (nbits & 63) p[nw - 1] &= (1ull << (nbits & 63)) - 1;
    return p;
}
/* out-of-range select reads x (IEEE 1364-2005 5.2.1): two-state stand-in 0 */
static inline uint64_t vs_wide_bit(const uint64_t *v, uint64_t nbits, uint64_t idx) {
    return idx < nbits ? (v[idx >> 6] >> (idx & 63)) & 1ull : 0ull;
}
static inline uint64_t vs_wide_get(const uint64_t *v, uint64_t nbits, uint64_t lsb, unsigned w) {
    if (lsb + w <= nbits) {                             /* fully inside the vector: at most two words */
        const uint64_t wi = lsb >> 6, sh = lsb & 63;
        uint64_t r = v[wi] >> sh;
        if (sh && sh + w > 64) r |= v[wi + 1] << (64 - sh);
        return w >= 64 ? r : r & ((1ull << w) - 1);
    }
    uint64_t r = 0;
    for (unsigned k = 0; k < w; k++) r |= vs_wide_bit(v, nbits, lsb + k) << k;
    return r;
}
static inline void vs_wide_load(uint64_t *v, uint64_t nbits, const uint64_t *words, uint64_t src_bits, int xfill) {
    /* a narrower source leaves the upper bits undriven (z): filled like every other x/z */
    for (uint64_t i = 0; i < nbits; i++) {
        const uint64_t bit = i < src_bits ? (words[i >> 6] >> (i & 63)) & 1ull : (uint64_t)(xfill != 0);
        v[i >> 6] = (v[i >> 6] & ~(1ull << (i & 63))) | (bit << (i & 63));
    }
}
static inline uint64_t vs_setbit(uint64_t old, uint64_t idx, unsigned width, uint64_t value) {
    if (idx >= width) return old;                       /* out-of-range write has no effect */
    return (old & ~(1ull << idx)) | ((value & 1ull) << idx);
}
static inline uint64_t vs_mem_rd(const uint64_t *mem, unsigned depth, uint64_t idx, int xfill, uint64_t mask) {
    return idx < depth ? mem[idx] : (xfill ? mask : 0ull);
}
static inline void vs_mem_wr(uint64_t *mem, unsigned depth, uint64_t idx, uint64_t value) {
    if (idx < depth) mem[idx] = value;
}

static inline vs_upd *vs_q_push(vs_model *m) {
    vs_header *h = VS_H(m);
    if (h->qn == h->qcap) {
        h->qcap = h->qcap ? 2 * h->qcap : 16;
        h->q = (vs_upd *)realloc(h->q, h->qcap * sizeof(vs_upd));
    }
    return &h->q[h->qn++];
}
static inline void vs_q_bit(vs_model *m, int dst, uint64_t idx, uint64_t value) {
    vs_upd *u = vs_q_push(m);
    u->kind = 0; u->dst = dst; u->a = (int64_t)idx; u->value = value & 1ull;
    if (idx > (uint64_t)INT64_MAX) u->a = INT64_MAX;
}
static inline void vs_q_part(vs_model *m, int dst, int64_t msb, int64_t lsb, uint64_t value) {
    vs_upd *u = vs_q_push(m);
    u->kind = 1; u->dst = dst; u->a = msb; u->b = lsb; u->value = value;
}
static inline void vs_q_copy(vs_model *m, int dst, int src) {
    vs_upd *u = vs_q_push(m);
    u->kind = 2; u->dst = dst; u->src = src;
}
/* Non-blocking commit: every source is read before any destination changes, then the updates are applied in
 * program order (IEEE 1364-2005 11.4: the NBA region executes the updates in the order they were scheduled). */
static inline void vs_q_commit(vs_model *m) {
    vs_header *h = VS_H(m);
    if (!h->qn) return;
    h->snap_n = 0;
    for (size_t k = 0; k < h->qn; k++) {
        vs_upd *u = &h->q[k];
        if (u->kind != 2) continue;
        const size_t nw = (size_t)((h->wide_w[u->src] + 63) / 64);
        if (h->snap_n + nw > h->snap_cap) {
            h->snap_cap = 2 * (h->snap_n + nw);
            h->snap = (uint64_t *)realloc(h->snap, h->snap_cap * 8);
        }
        memcpy(h->snap + h->snap_n, h->wide[u->src], nw * 8);
        u->snap_off = h->snap_n;
        h->snap_n += nw;
    }
    for (size_t k = 0; k < h->qn; k++) {
        const vs_upd *u = &h->q[k];
        uint64_t *d = h->wide[u->dst];
        const uint64_t dw = h->wide_w[u->dst];
        if (u->kind == 0) {
            if (u->a >= 0 && (uint64_t)u->a < dw) d[u->a >> 6] = (d[u->a >> 6] & ~(1ull << (u->a & 63))) | (u->value << (u->a & 63));
        } else if (u->kind == 1) {
            /* bits [msb:lsb] <= value: bit lsb + k takes bit k of the (zero-extended) 64-bit value */
            if (u->b == 0 && u->a >= 0 && (uint64_t)u->a + 1 >= dw && u->value == 0) { memset(d, 0, (size_t)((dw + 63) / 64) * 8); continue; }
            for (int64_t i = u->b < 0 ? 0 : u->b; i <= u->a && (uint64_t)i < dw; i++) {
                const int64_t kbit = i - u->b;
                const uint64_t bit = kbit < 64 ? (u->value >> kbit) & 1ull : 0ull;
                d[i >> 6] = (d[i >> 6] & ~(1ull << (i & 63))) | (bit << (i & 63));
            }
        } else {
            const uint64_t sw = h->wide_w[u->src];
            const uint64_t *s = h->snap + u->snap_off;
            if (sw == dw) { memcpy(d, s, (size_t)((dw + 63) / 64) * 8); continue; }
            for (uint64_t i = 0; i < dw; i++) {        /* truncation / zero-extension to the destination width */
                const uint64_t bit = i < sw ? (s[i >> 6] >> (i & 63)) & 1ull : 0ull;
                d[i >> 6] = (d[i >> 6] & ~(1ull << (i & 63))) | (bit << (i & 63));
            }
        }
    }
    h->qn = 0;
}

/* ---- interface of a generated model ---- */
const char *vs_module_name(void);
vs_model *vs_new(const char *const *param_names, const uint64_t *param_values, int n_params, int xfill);
void vs_delete(vs_model *m);
int vs_set(vs_model *m, const char *port, uint64_t value);                                   /* scalar input port */
int vs_set_wide(vs_model *m, const char *port, const uint64_t *words, uint64_t nbits);       /* wide input port */
int vs_get(const vs_model *m, const char *name, uint64_t *value);                            /* any scalar reg */
uint64_t *vs_ptr(vs_model *m, const char *name);        /* storage of a scalar's committed value (stable for the model's life) */
uint64_t *vs_get_wide(vs_model *m, const char *name, uint64_t *nbits);
void vs_eval_comb(vs_model *m);
void vs_posedge(vs_model *m);
#endif
