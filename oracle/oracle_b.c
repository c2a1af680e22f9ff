/*
 * oracle_b.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 * Functional restatement of what Design/FPGA.v computes per stream, with the testbench's
 * observation rules (Simulation/testbench_BLK_Mem.sv) folded in:
 *   - start set {0}, set only at reset, no re-injection            Design/FPGA.v:134-147
 *   - a state is expanded when active; every CSR entry of its row whose symbol equals the
 *     stream's current byte sets the target bit in `next`           Design/FPGA.v:264-268
 *   - row of state s = entries E[size+1+row_ptr[s] .. size+1+row_ptr[s+1])   FPGA.v:773,782,793
 *   - entry = {symbol[31:24], target[23:0]}                          Design/FPGA.v:888-898
 *   - zero out-degree == accepting; pulse while scanning state i     Design/FPGA.v:210-226
 *   - states are visited in ascending i within a step                Design/FPGA.v:725-728,747-751
 *   - current <= next; next <= 0 after the last state                Design/FPGA.v:730-741,753-764
 */
#include "oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

static int cmp_u32(const void *a, const void *b) {
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return (x > y) - (x < y);
}

int orc_b_scan(const uint32_t *E, size_t n_entries, uint32_t size, const uint8_t *data,
               uint64_t n_steps, uint32_t stream_id, uint64_t *counts, orc_rec *recs,
               uint64_t cap, uint64_t *n_recs, uint64_t *sum_active, uint32_t *max_active) {
    if (size == 0 || (size_t)size + 1 > n_entries) return -1;
    const uint32_t *rp = E;
    const uint32_t *tr = E + size + 1;
    if ((size_t)size + 1 + rp[size] > n_entries) return -1;
    uint32_t *cur = (uint32_t *)malloc(sizeof(uint32_t) * size);
    uint32_t *nxt = (uint32_t *)malloc(sizeof(uint32_t) * size);
    uint64_t *stamp = (uint64_t *)calloc(size, sizeof(uint64_t)); /* step+1 at which state was added */
    uint32_t ncur = 1, nnxt;
    uint64_t nr = 0, sa = 0;
    uint32_t ma = 0;
    cur[0] = 0;
    for (uint64_t k = 0; k < n_steps; k++) {
        uint8_t c = data[k];
        if (ncur > 1) qsort(cur, ncur, sizeof(uint32_t), cmp_u32);
        sa += ncur;
        if (ncur > ma) ma = ncur;
        nnxt = 0;
        for (uint32_t a = 0; a < ncur; a++) {
            uint32_t s = cur[a];
            uint32_t b = rp[s], e = rp[s + 1];
            if (b == e) { /* accepting */
                if (counts) counts[s]++;
                if (recs && nr < cap) { recs[nr].stream = stream_id; recs[nr].pos = (uint32_t)k; recs[nr].state = s; }
                nr++;
                continue;
            }
            for (uint32_t j = b; j < e; j++) {
                uint32_t w = tr[j];
                if ((w >> 24) == c) {
                    uint32_t t = w & 0xFFFFFFu;
                    if (t >= size) { free(cur); free(nxt); free(stamp); return -2; }
                    if (stamp[t] != k + 1) { stamp[t] = k + 1; nxt[nnxt++] = t; }
                }
            }
        }
        uint32_t *tmp = cur; cur = nxt; nxt = tmp;
        ncur = nnxt;
    }
    free(cur); free(nxt); free(stamp);
    if (n_recs) *n_recs = nr;
    if (sum_active) *sum_active = sa;
    if (max_active) *max_active = ma;
    return 0;
}

typedef struct {
    const uint32_t *E; size_t n_entries; uint32_t size; const uint8_t *data;
    uint64_t s0, s1, stride, n_steps;
    uint64_t *counts; orc_rec *recs; uint64_t rcap, nrec, sum_active; int rc; int want_recs;
} bjob;

static void *bworker(void *arg) {
    bjob *j = (bjob *)arg;
    j->counts = (uint64_t *)calloc(j->size, sizeof(uint64_t));
    j->rcap = j->want_recs ? 1024 : 0;
    j->recs = j->want_recs ? (orc_rec *)malloc(j->rcap * sizeof(orc_rec)) : NULL;
    j->nrec = 0; j->sum_active = 0; j->rc = 0;
    for (uint64_t s = j->s0; s < j->s1; s++) {
        for (;;) {
            uint64_t nr = 0, sa = 0;
            /* counts are only committed once the record buffer was large enough */
            uint64_t *tmpc = (uint64_t *)calloc(j->size, sizeof(uint64_t));
            int rc = orc_b_scan(j->E, j->n_entries, j->size, j->data + s * j->stride, j->n_steps,
                                (uint32_t)s, tmpc, j->recs ? j->recs + j->nrec : NULL,
                                j->rcap - j->nrec, &nr, &sa, NULL);
            if (rc) { j->rc = rc; free(tmpc); return NULL; }
            if (j->want_recs && nr > j->rcap - j->nrec) {
                j->rcap = (j->nrec + nr) * 2;
                j->recs = (orc_rec *)realloc(j->recs, j->rcap * sizeof(orc_rec));
                free(tmpc);
                continue;
            }
            for (uint32_t q = 0; q < j->size; q++) j->counts[q] += tmpc[q];
            free(tmpc);
            j->nrec += nr; j->sum_active += sa;
            break;
        }
    }
    return NULL;
}

int orc_b_scan_many(const uint32_t *E, size_t n_entries, uint32_t size, const uint8_t *data,
                    uint64_t n_streams, uint64_t stride, uint64_t n_steps, int n_threads,
                    uint64_t *counts, orc_rec *recs, uint64_t cap, uint64_t *n_recs,
                    uint64_t *sum_active) {
    if (n_threads < 1) n_threads = 1;
    if ((uint64_t)n_threads > n_streams && n_streams > 0) n_threads = (int)n_streams;
    bjob *jobs = (bjob *)calloc((size_t)n_threads, sizeof(bjob));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; t++) {
        jobs[t].E = E; jobs[t].n_entries = n_entries; jobs[t].size = size; jobs[t].data = data;
        jobs[t].stride = stride; jobs[t].n_steps = n_steps; jobs[t].want_recs = recs != NULL;
        jobs[t].s0 = n_streams * (uint64_t)t / (uint64_t)n_threads;
        jobs[t].s1 = n_streams * (uint64_t)(t + 1) / (uint64_t)n_threads;
        pthread_create(&th[t], NULL, bworker, &jobs[t]);
    }
    uint64_t nr = 0, sa = 0;
    int rc = 0;
    for (int t = 0; t < n_threads; t++) {
        pthread_join(th[t], NULL);
        if (jobs[t].rc) rc = jobs[t].rc;
        if (!jobs[t].rc) {
            if (counts) for (uint32_t q = 0; q < size; q++) counts[q] += jobs[t].counts[q];
            for (uint64_t r = 0; r < jobs[t].nrec; r++, nr++)
                if (recs && nr < cap) recs[nr] = jobs[t].recs[r];
            if (!recs) nr += 0;
            sa += jobs[t].sum_active;
        }
        free(jobs[t].counts); free(jobs[t].recs);
    }
    if (!recs) { /* record total still reported */
        nr = 0;
        for (int t = 0; t < n_threads; t++) nr += jobs[t].nrec;
    }
    free(jobs); free(th);
    if (n_recs) *n_recs = nr;
    if (sum_active) *sum_active = sa;
    return rc;
}
