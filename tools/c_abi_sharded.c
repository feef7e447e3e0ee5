/* c_abi_sharded.c -- the multi-GPU entry points driven from plain C, no Python, no torch, no CUDA
 * API of its own: what a Rust host binding include/cnvetti_b200.h would do.
 *
 *   gcc -O2 -I include tools/c_abi_sharded.c -o /tmp/c_abi_sharded -L cnvetti_b200 -lcnvetti_b200 \
 *       -Wl,-rpath,$PWD/cnvetti_b200 -lpthread -lm
 *   /tmp/c_abi_sharded [world]        (world = number of GPUs to use, default: 2; 1 works without NCCL)
 *
 * One host thread per GPU (the reference runs one `run` per rayon worker,
 * cnvetti/src/quick_wis_build_model.rs:148-184).  Each thread makes its ctx + communicator, passes ITS
 * rows of a synthetic panel to cnvb_wis_build_sharded and ITS records to the two sharded normalisers;
 * thread 0 also computes the single-GPU results and everything must agree bit for bit.  Exit code 0 = ok. */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cnvetti_b200.h"

enum { T = 6000, S = 48, K = 40, NREC = 300000 };

static float *X;
static uint32_t *tid;
static float *lcv, *gc;
static int world = 2;
static cnvb_comm_id comm_id;
static pthread_barrier_t bar;

typedef struct {
    int rank, rc;
    float thr;
    float *dist;
    uint32_t *idx, *len, *calls;
    uint8_t *filt;
    float *cv_tot, *cv_gc, *cv2;
    uint8_t *flags;
    float med[101];
    double sum;
} rank_out;
static rank_out outs[16];

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static double rnd(void) { /* xorshift64*, uniform in [0, 1) */
    rng_state ^= rng_state >> 12;
    rng_state ^= rng_state << 25;
    rng_state ^= rng_state >> 27;
    return (double)((rng_state * 0x2545F4914F6CDD1Dull) >> 11) / 9007199254740992.0;
}
static double gauss(void) { return sqrt(-2.0 * log(rnd() + 1e-300)) * cos(6.283185307179586 * rnd()); }

static void split(uint64_t n, int r, uint64_t *lo, uint64_t *hi) { /* uneven on purpose */
    uint64_t cut[17];
    for (int q = 0; q <= world; ++q) cut[q] = n * (uint64_t)q / (uint64_t)world;
    for (int q = 1; q < world; ++q) cut[q] += (uint64_t)(q % 2 ? 37 : -11);
    *lo = cut[r];
    *hi = cut[r + 1];
}

#define CHECK(call)                                                                         \
    do {                                                                                    \
        int rc__ = (call);                                                                  \
        if (rc__ != 0) {                                                                    \
            fprintf(stderr, "rank %d: %s failed (%d): %s\n", o->rank, #call, rc__,          \
                    ctx ? cnvb_last_error(ctx) : cnvb_create_error());                      \
            o->rc = rc__;                                                                   \
            return NULL;                                                                    \
        }                                                                                   \
    } while (0)

static void *worker(void *arg) {
    rank_out *o = (rank_out *)arg;
    cnvb_ctx *ctx = NULL;
    cnvb_comm *comm = NULL;
    CHECK(cnvb_ctx_create(o->rank, NULL, &ctx));
    if (world > 1) {
        if (o->rank == 0 && cnvb_comm_unique_id(&comm_id) != 0) {
            fprintf(stderr, "no NCCL: %s\n", cnvb_comm_create_error());
            o->rc = CNVB_ERR_COMM;
        }
        pthread_barrier_wait(&bar); /* the 128 bytes reach the other "processes" */
        if (outs[0].rc) return NULL;
        CHECK(cnvb_comm_init_rank(ctx, &comm_id, o->rank, world, &comm));
    } else {
        CHECK(cnvb_comm_from_nccl(ctx, NULL, 0, 1, &comm));
    }
    cnvb_wis_opts wo = {5.64, 0.35, 10, K, 4, CNVB_WIS_ENGINE_TENSOR};
    uint64_t lo, hi;
    split(T, o->rank, &lo, &hi);
    const uint32_t rows = (uint32_t)(hi - lo);
    o->dist = malloc((size_t)rows * K * 4);
    o->idx = malloc((size_t)rows * K * 4);
    o->len = malloc((size_t)rows * 4);
    o->calls = malloc((size_t)rows * 4);
    o->filt = malloc(rows);
    uint32_t row_begin = 0, t_all = 0;
    CHECK(cnvb_wis_build_sharded(comm, X + lo * S, tid + lo, rows, S, &wo, o->dist, o->idx, o->len, o->calls, o->filt,
                                 &o->thr, &row_begin, &t_all, CNVB_MEM_HOST));
    if (row_begin != lo || t_all != T) {
        fprintf(stderr, "rank %d: row block [%u, ..) of %u targets, expected [%llu, ..) of %d\n", o->rank, row_begin, t_all,
                (unsigned long long)lo, T);
        o->rc = 1;
        return NULL;
    }
    uint64_t a, b;
    split(NREC, o->rank, &a, &b);
    const uint64_t n = b - a;
    o->cv_tot = malloc(n * 4);
    o->cv_gc = malloc(n * 4);
    o->cv2 = malloc(n * 4);
    o->flags = malloc(n);
    CHECK(cnvb_norm_total_sharded(comm, lcv + a, NULL, n, o->cv_tot, NULL, &o->sum, CNVB_MEM_HOST));
    CHECK(cnvb_norm_gc_median_sharded(comm, lcv + a, NULL, gc + a, n, o->cv_gc, o->cv2, NULL, o->flags, o->med, CNVB_MEM_HOST));
    cnvb_comm_destroy(comm);
    if (o->rank == 0) { /* the single-GPU results, on the same device */
        float *dist = malloc((size_t)T * K * 4), thr = 0, *cvt = malloc(NREC * 4), *cvg = malloc(NREC * 4), *cv2 = malloc(NREC * 4);
        uint32_t *idx = malloc((size_t)T * K * 4), *len = malloc(T * 4), *calls = malloc(T * 4);
        uint8_t *filt = malloc(T), *flags = malloc(NREC);
        float med[101];
        double sum = 0;
        CHECK(cnvb_wis_build(ctx, X, tid, T, S, &wo, dist, idx, len, calls, filt, &thr, CNVB_MEM_HOST));
        CHECK(cnvb_norm_total(ctx, lcv, NULL, NREC, cvt, NULL, &sum, CNVB_MEM_HOST));
        CHECK(cnvb_norm_gc_median(ctx, lcv, NULL, gc, NREC, cvg, cv2, NULL, flags, med, CNVB_MEM_HOST));
        pthread_barrier_wait(&bar); /* all ranks have their results */
        int bad = 0;
        for (int r = 0; r < world; ++r) {
            rank_out *q = &outs[r];
            if (q->rc) return NULL;
            uint64_t l, h, ra, rb;
            split(T, r, &l, &h);
            split(NREC, r, &ra, &rb);
            bad |= memcmp(&q->thr, &thr, 4) != 0;
            bad |= memcmp(q->dist, dist + l * K, (h - l) * K * 4) != 0;
            bad |= memcmp(q->len, len + l, (h - l) * 4) != 0;
            bad |= memcmp(q->calls, calls + l, (h - l) * 4) != 0;
            bad |= memcmp(q->filt, filt + l, h - l) != 0;
            for (uint64_t t = l; t < h; ++t) /* ref_idx is defined up to ref_len entries after the prune */
                bad |= memcmp(q->idx + (t - l) * K, idx + t * K, (size_t)len[t] * 4) != 0;
            bad |= memcmp(&q->sum, &sum, 8) != 0;
            bad |= memcmp(q->cv_tot, cvt + ra, (rb - ra) * 4) != 0;
            bad |= memcmp(q->med, med, sizeof(med)) != 0;
            bad |= memcmp(q->cv_gc, cvg + ra, (rb - ra) * 4) != 0;
            bad |= memcmp(q->cv2, cv2 + ra, (rb - ra) * 4) != 0;
            bad |= memcmp(q->flags, flags + ra, rb - ra) != 0;
            if (bad) {
                fprintf(stderr, "rank %d: sharded result differs from the single-GPU result\n", r);
                break;
            }
        }
        o->rc = bad;
        if (!bad) printf("c_abi_sharded ok: world %d, %d targets x %d samples, threshold %.9g, %d records, sum %.17g\n", world, T,
                         S, thr, NREC, sum);
    } else {
        pthread_barrier_wait(&bar);
    }
    cnvb_ctx_destroy(ctx);
    return NULL;
}

int main(int argc, char **argv) {
    if (argc > 1) world = atoi(argv[1]);
    if (world < 1 || world > 16) return 2;
    X = malloc((size_t)T * S * 4);
    tid = malloc(T * 4);
    lcv = malloc(NREC * 4);
    gc = malloc(NREC * 4);
    for (int t = 0; t < T; ++t) {
        const double eff = exp(0.4 * gauss());
        tid[t] = (uint32_t)(t * 7 / T);
        for (int s = 0; s < S; ++s) X[(size_t)t * S + s] = (float)fmax(0.0, eff * (1.0 + 0.08 * gauss()) / T);
    }
    for (int i = 0; i < NREC; ++i) {
        gc[i] = (float)(0.3 + 0.25 * rnd());
        lcv[i] = (float)((double)(int)(100.0 + 30.0 * gauss()) / 500.0);
        if (i % 1000 == 7) gc[i] = NAN;         /* window without ACGT */
        if (i % 50000 == 9) gc[i] = 0.97f;      /* a GC bin with fewer than 10 windows */
    }
    pthread_barrier_init(&bar, NULL, (unsigned)world);
    pthread_t th[16];
    for (int r = 0; r < world; ++r) {
        outs[r].rank = r;
        pthread_create(&th[r], NULL, worker, &outs[r]);
    }
    int rc = 0;
    for (int r = 0; r < world; ++r) {
        pthread_join(th[r], NULL);
        rc |= outs[r].rc;
    }
    return rc ? 1 : 0;
}
