/*
 * oracle_a.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Cycle-level restatement of the reference hardware, one C statement group per Verilog branch:
 *   CSR_traversal sequential block      Design/FPGA.v:115-768
 *   address generator (combinational)   Design/FPGA.v:771-874
 *   bus unpack (combinational)          Design/FPGA.v:876-900
 *   design_1_wrapper                    Design/top.v:10-13 -- source absent from the reference;
 *       modelled as a synchronous ROM with one cycle of latency (dout <= mem[addr] every edge),
 *       the only latency consistent with FPGA.v:161->182 and :233->259-305
 *   testbench feeder/counters/finish    Simulation/testbench_BLK_Mem.sv:49-86
 *
 * Registers use non-blocking semantics: every right-hand side below reads the pre-edge copy `r`,
 * every assignment writes the post-edge copy `n`.  Bits set in next/next_2 are applied in place,
 * which is equivalent because no branch reads next/next_2 in the same edge in which it sets bits.
 * Registers the reference declares but never reads for results (first, accepting, range_last,
 * flag_check, check, ...; FPGA.v:45,52,59,66,73-74,98,113) are not modelled.
 */
#include "oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint32_t i;              /* FPGA.v:41  reg [19:0] */
    uint32_t state;          /* FPGA.v:71  reg [2:0]  */
    uint32_t flag;           /* FPGA.v:51  */
    uint32_t flag_2;         /* FPGA.v:111 */
    uint32_t flag_1_or_2;    /* FPGA.v:104 */
    uint32_t range;          /* FPGA.v:65  reg [23:0] */
    uint32_t up_counter;     /* FPGA.v:68  reg [23:0] */
    uint32_t rd_address;     /* FPGA.v:39  reg [15:0] */
    uint32_t bo_reg;         /* block_offset_reg, FPGA.v:103 */
    uint32_t bo1_reg;        /* block_offset_plus_one_reg, FPGA.v:102 */
    uint32_t cache_temp;     /* FPGA.v:95 */
    uint32_t range_next;     /* FPGA.v:96 */
    uint32_t bo_f0;          /* block_offset_flag_0, FPGA.v:100 */
    uint32_t ncb_f0, ncb_f1, ncb_f2, ncb_f2_prev; /* no_cached_blocks_flag_*, FPGA.v:107-110 */
    uint32_t r1s, r2s;       /* range_1_state / range_2_state, FPGA.v:76-77 */
    uint32_t icf;            /* input_char_flag */
    uint32_t amf, amf2;      /* accepting_match_flag(_2) */
    uint32_t dout[4];        /* BRAM output register; dout[k] = cache[k] (FPGA.v:881-884) */
} regs_t;

typedef struct {
    uint32_t size;
    uint32_t addr_mask;
    const uint32_t *E;
    size_t n_entries;
    uint64_t nw;             /* bitmap words */
    uint64_t *cur, *cur2, *nxt, *nxt2;
    regs_t r;
} sim_t;

static inline int getbit(const uint64_t *b, uint32_t i) { return (int)((b[i >> 6] >> (i & 63)) & 1u); }
static inline void setbit(uint64_t *b, uint32_t i) { b[i >> 6] |= (uint64_t)1 << (i & 63); }

/* ROM read of one 128-bit line; lines past the image read as zero. */
static inline void rom_read(const sim_t *s, uint32_t line, uint32_t out[4]) {
    size_t e = (size_t)line * 4;
    for (int k = 0; k < 4; k++) out[k] = (e + (size_t)k < s->n_entries) ? s->E[e + (size_t)k] : 0u;
}

/* "compare lane k" -- FPGA.v:264-268 and its 27 textual repetitions. */
#define LANE(k)                                                                      \
    do {                                                                             \
        uint32_t w_ = r.dout[(k)], sy_ = w_ >> 24, tg_ = w_ & 0xFFFFFFu;             \
        if (ci && sy_ == ch1 && tg_ < s->size) setbit(s->nxt, tg_);                  \
        if (ci2 && sy_ == ch2 && tg_ < s->size) setbit(s->nxt2, tg_);                \
    } while (0)

/* lanes valid for the line fetched in the flag==0 state: FPGA.v:262-305 (and :421-464, :564-607) */
#define LANES_MASK0()                                                                \
    do {                                                                             \
        for (uint32_t k_ = 0; k_ < 4; k_++)                                          \
            if (k_ >= r.bo_f0 && r.ncb_f0 >= k_ - r.bo_f0 + 1) LANE(k_);             \
    } while (0)
/* lanes < n: FPGA.v:315-349 and repetitions */
#define LANES_LT(nv)                                                                 \
    do {                                                                             \
        for (uint32_t k_ = 0; k_ < 4; k_++) if ((nv) > k_) LANE(k_);                 \
    } while (0)

static void swap_sets(sim_t *s) { /* FPGA.v:733-737 / :756-760 */
    uint64_t *t;
    t = s->cur; s->cur = s->nxt; s->nxt = t;
    t = s->cur2; s->cur2 = s->nxt2; s->nxt2 = t;
    memset(s->nxt, 0, s->nw * 8);
    memset(s->nxt2, 0, s->nw * 8);
}

/* One posedge clk with reset == 0. */
static void edge(sim_t *s, uint32_t ch1, uint32_t ch2) {
    const regs_t r = s->r;
    regs_t n = r;
    const uint32_t size = s->size;
    const int ci = getbit(s->cur, r.i), ci2 = getbit(s->cur2, r.i);
    const int active = ci || ci2;

    /* ---- combinational address generator, FPGA.v:771-874 ---- */
    uint32_t block_offset = 0, block_offset_p1 = 0, cache_line_no = 0;
    uint32_t ncb = 0, up_int = 0, range_int = 0;
    const uint32_t offset = (size + 1) & 0x1FFFFFFu;                 /* :773, reg [24:0] */
    if (active && r.state == 0) {                                     /* :780-786 */
        uint32_t rai = r.i;
        block_offset = rai & 3u;
        block_offset_p1 = block_offset + 1u;
        cache_line_no = (rai >> 2) & s->addr_mask;
    }
    if (r.flag == 0 && r.state == 3 && r.range > 0) {                 /* :788-817 */
        uint32_t rai = (offset + r.up_counter) & 0x1FFFFFFu;
        block_offset = rai & 3u;
        cache_line_no = (rai >> 2) & s->addr_mask;
        uint32_t ncb_int = 4u - block_offset;
        ncb = (r.range > ncb_int) ? ncb_int : r.range;
        up_int = ((r.range > ncb) ? r.up_counter + ncb : r.up_counter + r.range) & 0xFFFFFFu;
        range_int = (r.range > ncb) ? r.range - ncb : 0u;
    } else if ((r.flag == 1 || r.flag == 2) && r.state == 3 && r.range > 0) { /* :818-867 */
        cache_line_no = (r.rd_address + 1u) & s->addr_mask;
        ncb = (r.range > 4u) ? 4u : r.range;
        up_int = ((r.range > 4u) ? r.up_counter + 4u : r.up_counter + r.range) & 0xFFFFFFu;
        range_int = (r.range > 4u) ? r.range - 4u : 0u;
    }

    /* ---- sequential block, FPGA.v:155-767 (priority if / else-if chain) ---- */
    if (active && r.state == 0) {                                     /* :158-165 */
        n.icf = 0;
        n.rd_address = cache_line_no;
        n.bo_reg = block_offset;
        n.bo1_reg = block_offset_p1;
        n.state = 1;
    } else if (r.state == 1) {                                        /* :166-175 */
        if (r.bo_reg == 3) n.rd_address = (r.rd_address + 1u) & s->addr_mask;
        n.state = 2;
    } else if (r.state == 2) {                                        /* :176-207 */
        if (r.bo_reg != 3) {
            n.range = (r.dout[r.bo1_reg & 3u] - r.dout[r.bo_reg]) & 0xFFFFFFu;
            n.up_counter = r.dout[r.bo_reg] & 0xFFFFFFu;
            n.flag = 0;
            n.state = 3;
        } else if (r.range_next == 0) {
            n.range_next = 1;
            n.cache_temp = r.dout[3];
        } else {
            n.range_next = 0;
            n.range = (r.dout[0] - r.cache_temp) & 0xFFFFFFu;
            n.up_counter = r.cache_temp & 0xFFFFFFu;
            n.flag = 0;
            n.state = 3;
        }
    } else if (r.state == 3) {                                        /* :208-716 */
        if (r.range == 0 && r.flag == 0) {                            /* :210-226 accepting */
            if (ci) n.amf = 1;
            if (ci2) n.amf2 = 1;
            n.state = 4;
        } else if (r.range > 0) {                                     /* :227-407 */
            if (r.flag == 0) {                                        /* :229-242 */
                n.rd_address = cache_line_no;
                n.flag_1_or_2 = 0;
                n.bo_f0 = block_offset;
                n.ncb_f0 = ncb;
                n.range = range_int;
                n.up_counter = up_int;
                n.flag = 1;
            } else if (r.flag == 1) {                                 /* :243-254 */
                n.flag = 2;
                n.rd_address = cache_line_no;
                n.flag_1_or_2 = 1;
                n.ncb_f1 = ncb;
                n.range = range_int;
                n.up_counter = up_int;
                n.flag_2 = 0;
            } else if (r.flag == 2) {                                 /* :255-406 */
                if (r.flag_2 == 0) { LANES_MASK0(); n.flag_2 = 1; }   /* :259-310 */
                else if (r.flag_2 <= 1) { LANES_LT(r.ncb_f1); n.flag_2 = 2; } /* :312-354 */
                else if (r.flag_2 <= 2) { LANES_LT(r.ncb_f2); }       /* :356-396 */
                n.flag_1_or_2 = 2;                                    /* :398-405 */
                n.ncb_f2_prev = r.ncb_f2;
                n.ncb_f2 = ncb;
                n.range = range_int;
                n.up_counter = up_int;
                n.rd_address = cache_line_no;
            }
        } else { /* range == 0 with lines still in flight, :408-714 */
            if (r.flag == 1 && r.r1s == 0) {                          /* :411-414 */
                n.r1s = 1;
            } else if (r.flag == 1 && r.r1s == 1) {                   /* :415-473 */
                if (r.flag_1_or_2 == 0) LANES_MASK0();
                n.r1s = 0; n.state = 4; n.flag = 0;
            } else if (r.flag == 2 && r.r2s == 0) {                   /* :474-613 */
                if (r.flag_2 == 2) LANES_LT(r.ncb_f2_prev);
                if (r.flag_2 == 1) LANES_LT(r.ncb_f1);
                if (r.flag_1_or_2 == 1) LANES_MASK0();
                n.r2s = 1;
            } else if (r.flag == 2 && r.r2s == 1) {                   /* :614-706 */
                if (r.flag_1_or_2 == 1) LANES_LT(r.ncb_f1);
                if (r.flag_1_or_2 == 2) LANES_LT(r.ncb_f2);
                n.r2s = 0; n.state = 4; n.flag = 0;
            } else {                                                  /* :707-712 */
                n.state = 4; n.flag = 0;
            }
        }
    } else if (r.state == 4) {                                        /* :717-743 */
        n.amf = 0; n.amf2 = 0; n.flag_2 = 0; n.state = 0;
        if (r.i < size - 1) n.i = r.i + 1;
        else { swap_sets(s); n.i = 0; n.icf = 1; }
    } else if (!ci && !ci2 && r.state == 0) {                         /* :744-765 */
        if (r.i < size - 1) { n.icf = 0; n.i = r.i + 1; }
        else { swap_sets(s); n.i = 0; n.icf = 1; }
    }

    /* ---- ROM: address registered before this edge is answered at this edge ---- */
    rom_read(s, r.rd_address, n.dout);
    s->r = n;
}

static uint32_t next_active(const sim_t *s, uint32_t from, uint32_t limit) {
    /* smallest j in [from, limit] with cur[j] | cur2[j], else limit */
    for (uint32_t j = from; j <= limit;) {
        uint64_t w = (s->cur[j >> 6] | s->cur2[j >> 6]) >> (j & 63);
        if (w) {
            uint32_t q = j + (uint32_t)__builtin_ctzll(w);
            return q < limit ? q : limit;
        }
        j = (j | 63u) + 1u;
    }
    return limit;
}

static int a_run(const uint32_t *E, size_t n_entries, uint32_t size, const uint8_t *lo,
              const uint8_t *hi, uint64_t M, int addr_bits, int fast_idle, uint16_t *mc1,
              uint16_t *mc2, uint64_t *cnt1, uint64_t *cnt2, orc_rec *recs, uint64_t cap,
              uint64_t *n_recs, uint64_t *cycles_out, uint64_t *trace, uint64_t trace_cap) {
    if (size == 0 || (size_t)size + 1 > n_entries || addr_bits < 1 || addr_bits > 30) return -1;
    sim_t s;
    memset(&s, 0, sizeof s);
    s.size = size; s.E = E; s.n_entries = n_entries;
    s.addr_mask = (addr_bits >= 32) ? 0xFFFFFFFFu : ((1u << addr_bits) - 1u);
    s.nw = ((uint64_t)size + 63) / 64;
    s.cur = (uint64_t *)calloc(s.nw, 8); s.cur2 = (uint64_t *)calloc(s.nw, 8);
    s.nxt = (uint64_t *)calloc(s.nw, 8); s.nxt2 = (uint64_t *)calloc(s.nw, 8);
    if (mc1) memset(mc1, 0, sizeof(uint16_t) * size);                 /* TB:41-45 */
    if (mc2) memset(mc2, 0, sizeof(uint16_t) * size);

    /* edge #1: reset branch, FPGA.v:119-153 */
    s.r.i = 0; s.r.state = 0; s.r.icf = 1; s.r.amf = 0; s.r.amf2 = 0;
    s.r.r2s = 0; s.r.r1s = 0; s.r.range_next = 0; s.r.range = 0;
    setbit(s.cur, 0); setbit(s.cur2, 0);
    uint64_t cycles = 0, m = 0, nr = 0;
    uint32_t ch1 = 0, ch2 = 0;
    int first = 1;
    for (;;) {
        if (!first) {
            if (fast_idle && s.r.state == 0 && s.r.i < size - 1 &&
                !getbit(s.cur, s.r.i) && !getbit(s.cur2, s.r.i)) {
                /* run of FPGA.v:747-752 edges: icf<=0, i<=i+1, nothing else changes */
                uint32_t j = next_active(&s, s.r.i + 1, size - 1);
                cycles += j - s.r.i;
                s.r.i = j; s.r.icf = 0;
                rom_read(&s, s.r.rd_address, s.r.dout);
                /* TB sees icf==0, flags==0 on every skipped edge: nothing to do */
                continue;
            }
            edge(&s, ch1, ch2);
        }
        first = 0;
        /* ---- testbench, TB:49-86, evaluated on post-edge values ---- */
        cycles++;                                                     /* TB:52 */
        if (trace && cycles <= trace_cap)                             /* what the outside sees after this edge */
            trace[cycles - 1] = (uint64_t)s.r.i | ((uint64_t)s.r.icf << 20) | ((uint64_t)s.r.amf << 21) |
                                ((uint64_t)s.r.amf2 << 22) | ((uint64_t)s.r.state << 24) | ((uint64_t)s.r.rd_address << 32);
        if (s.r.icf) { ch1 = lo[m]; ch2 = hi[m]; m++; }               /* TB:53-59 */
        if (s.r.amf) {                                                /* TB:61-64 */
            if (mc1) mc1[s.r.i] = (uint16_t)((mc1[s.r.i] + 1u) & 0x3FFu);
            if (cnt1) cnt1[s.r.i]++;
            if (recs && nr < cap) { recs[nr].stream = 0; recs[nr].pos = (uint32_t)(m - 1); recs[nr].state = s.r.i; }
            nr++;
        }
        if (s.r.amf2) {                                               /* TB:66-69 */
            if (mc2) mc2[s.r.i] = (uint16_t)((mc2[s.r.i] + 1u) & 0x3FFu);
            if (cnt2) cnt2[s.r.i]++;
            if (recs && nr < cap) { recs[nr].stream = 1; recs[nr].pos = (uint32_t)(m - 1); recs[nr].state = s.r.i; }
            nr++;
        }
        if (m == M) break;                                            /* TB:71 */
    }
    free(s.cur); free(s.cur2); free(s.nxt); free(s.nxt2);
    if (n_recs) *n_recs = nr;
    if (cycles_out) *cycles_out = cycles;
    return 0;
}

int orc_a_run(const uint32_t *E, size_t n_entries, uint32_t size, const uint8_t *lo,
              const uint8_t *hi, uint64_t M, int addr_bits, int fast_idle, uint16_t *mc1,
              uint16_t *mc2, uint64_t *cnt1, uint64_t *cnt2, orc_rec *recs, uint64_t cap,
              uint64_t *n_recs, uint64_t *cycles_out) {
    return a_run(E, n_entries, size, lo, hi, M, addr_bits, fast_idle, mc1, mc2, cnt1, cnt2, recs, cap, n_recs, cycles_out, NULL, 0);
}

/* Per-edge view of the ports and of `state` (i | input_char_flag << 20 | accepting_match_flag << 21 |
 * accepting_match_flag_2 << 22 | state << 24 | rd_address << 32, sampled after every posedge, the reset edge
 * first), for the first trace_cap edges: compared edge by edge with the reference's own Verilog executed through
 * oracle/vsim (tests/test_ref_vsim.py). */
int orc_a_trace(const uint32_t *E, size_t n_entries, uint32_t size, const uint8_t *lo, const uint8_t *hi,
                uint64_t M, int addr_bits, uint64_t *trace, uint64_t trace_cap, uint64_t *cycles_out) {
    return a_run(E, n_entries, size, lo, hi, M, addr_bits, 0, NULL, NULL, NULL, NULL, NULL, 0, NULL, cycles_out, trace, trace_cap);
}

/* SURVEY Appendix B.3 closed form, evaluated from functional sets of both streams. */
int orc_cycle_model(const uint32_t *E, size_t n_entries, uint32_t size, const uint8_t *lo,
                    const uint8_t *hi, uint64_t M, uint64_t *cycles_out) {
    if (size == 0 || (size_t)size + 1 > n_entries) return -1;
    const uint32_t *rp = E, *tr = E + size + 1;
    uint64_t nw = ((uint64_t)size + 63) / 64;
    uint64_t *c1 = (uint64_t *)calloc(nw, 8), *c2 = (uint64_t *)calloc(nw, 8);
    uint64_t *n1 = (uint64_t *)calloc(nw, 8), *n2 = (uint64_t *)calloc(nw, 8);
    setbit(c1, 0); setbit(c2, 0);
    uint64_t total = 1;
    for (uint64_t k = 0; k + 1 < M; k++) {
        uint64_t step = 0;
        for (uint32_t s = 0; s < size; s++) {
            int a1 = getbit(c1, s), a2 = getbit(c2, s);
            if (!a1 && !a2) { step += 1; continue; }
            uint32_t deg = rp[s + 1] - rp[s];
            uint64_t s3 = deg == 0 ? 1 : ((((uint64_t)size + 1 + rp[s]) % 4 + deg + 3) / 4 + 2);
            step += 1 + 1 + ((s % 4 == 3) ? 2 : 1) + s3 + 1;
            for (uint32_t j = rp[s]; j < rp[s + 1]; j++) {
                uint32_t w = tr[j], t = w & 0xFFFFFFu;
                if (t >= size) continue;
                if (a1 && (w >> 24) == lo[k]) setbit(n1, t);
                if (a2 && (w >> 24) == hi[k]) setbit(n2, t);
            }
        }
        total += step;
        uint64_t *t;
        t = c1; c1 = n1; n1 = t; memset(n1, 0, nw * 8);
        t = c2; c2 = n2; n2 = t; memset(n2, 0, nw * 8);
    }
    free(c1); free(c2); free(n1); free(n2);
    *cycles_out = total;
    return 0;
}

typedef struct {
    const uint32_t *E; size_t n_entries; uint32_t size; const uint8_t *data;
    uint64_t p0, p1, stride, M; int fast_idle;
    uint64_t *counts; uint64_t cycles, symbols; int rc;
} ajob;

static void *aworker(void *arg) {
    ajob *j = (ajob *)arg;
    j->counts = (uint64_t *)calloc(j->size, sizeof(uint64_t));
    uint64_t *c2 = (uint64_t *)calloc(j->size, sizeof(uint64_t));
    j->cycles = 0; j->symbols = 0; j->rc = 0;
    for (uint64_t p = j->p0; p < j->p1; p++) {
        uint64_t cyc = 0;
        int rc = orc_a_run(j->E, j->n_entries, j->size, j->data + (2 * p) * j->stride,
                           j->data + (2 * p + 1) * j->stride, j->M, 30, j->fast_idle, NULL, NULL,
                           j->counts, c2, NULL, 0, NULL, &cyc);
        if (rc) { j->rc = rc; break; }
        j->cycles += cyc;
        j->symbols += 2 * (j->M - 1);
    }
    for (uint32_t q = 0; q < j->size; q++) j->counts[q] += c2[q];
    free(c2);
    return NULL;
}

int orc_a_run_many(const uint32_t *E, size_t n_entries, uint32_t size, const uint8_t *data,
                   uint64_t n_pairs, uint64_t stride, uint64_t M, int n_threads, int fast_idle,
                   uint64_t *counts, uint64_t *total_cycles, uint64_t *total_symbols) {
    if (n_threads < 1) n_threads = 1;
    if ((uint64_t)n_threads > n_pairs && n_pairs > 0) n_threads = (int)n_pairs;
    ajob *jobs = (ajob *)calloc((size_t)n_threads, sizeof(ajob));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; t++) {
        jobs[t].E = E; jobs[t].n_entries = n_entries; jobs[t].size = size; jobs[t].data = data;
        jobs[t].stride = stride; jobs[t].M = M; jobs[t].fast_idle = fast_idle;
        jobs[t].p0 = n_pairs * (uint64_t)t / (uint64_t)n_threads;
        jobs[t].p1 = n_pairs * (uint64_t)(t + 1) / (uint64_t)n_threads;
        pthread_create(&th[t], NULL, aworker, &jobs[t]);
    }
    uint64_t cyc = 0, sym = 0;
    int rc = 0;
    for (int t = 0; t < n_threads; t++) {
        pthread_join(th[t], NULL);
        if (jobs[t].rc) rc = jobs[t].rc;
        if (counts) for (uint32_t q = 0; q < size; q++) counts[q] += jobs[t].counts[q];
        cyc += jobs[t].cycles; sym += jobs[t].symbols;
        free(jobs[t].counts);
    }
    free(jobs); free(th);
    if (total_cycles) *total_cycles = cyc;
    if (total_symbols) *total_symbols = sym;
    return rc;
}
