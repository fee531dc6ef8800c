/* oracle_batch.c -- multi-threaded batch drivers over the oracle, used only to time
 * the CPU port beside the GPU numbers (bench.py cpu_baseline) and to bulk-check
 * batches in tests.  TEST INFRASTRUCTURE ONLY (see oracle_capi.h).
 * One problem per thread at a time, a shared atomic cursor; mirrors
 * oracle/ref_glue/ref_capi.cpp:ref_batch so both CPU arms are driven identically. */
#include "oracle_capi.h"

#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <time.h>

typedef struct {
    const double* costs; const int64_t* costOff; const int32_t* nL; const int32_t* nM;
    int64_t n, k; double cutoff;
    double* probs; const int64_t* probOff;
    int64_t* col4row; const int64_t* c4rOff; int64_t* row4col; const int64_t* r4cOff;
    double* gain; int32_t* nFound;
    atomic_llong next;
} Batch;

static double now_s(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

static void* batch_worker(void* p) {
    Batch* b = (Batch*)p;
    for (;;) {
        int64_t i = (int64_t)atomic_fetch_add(&b->next, 1);
        if (i >= b->n) break;
        const int64_t nL = b->nL[i], nM = b->nM[i];
        const double* c = b->costs + b->costOff[i];
        if (b->probs) orc_assignment_prob(c, nL, nM, b->k, b->probs + b->probOff[i]);
        if (b->gain)
            b->nFound[i] = (int32_t)orc_kbest2d_cutoff(b->k, nL + nM, nM, 0, c, b->col4row + b->c4rOff[i],
                                                       b->row4col + b->r4cOff[i], b->gain + i * b->k, b->cutoff);
    }
    return NULL;
}

double orc_batch(const double* costs, const int64_t* costOff, const int32_t* nL, const int32_t* nM,
                 int64_t n, int64_t k, double cutoff, int nThreads,
                 double* probs, const int64_t* probOff,
                 int64_t* col4row, const int64_t* c4rOff, int64_t* row4col, const int64_t* r4cOff,
                 double* gain, int32_t* nFound) {
    Batch b = {costs, costOff, nL, nM, n, k, cutoff, probs, probOff, col4row, c4rOff, row4col, r4cOff, gain, nFound, 0};
    if (nThreads < 1) nThreads = 1;
    pthread_t* th = (pthread_t*)malloc((size_t)nThreads * sizeof(pthread_t));
    double t0 = now_s();
    for (int t = 0; t < nThreads; t++) pthread_create(&th[t], NULL, batch_worker, &b);
    for (int t = 0; t < nThreads; t++) pthread_join(th[t], NULL);
    double dt = now_s() - t0;
    free(th);
    return dt;
}

typedef struct { const double* mats; int64_t dim, n; double* out; atomic_llong next; } PermBatch;

static void* perm_worker(void* p) {
    PermBatch* b = (PermBatch*)p;
    for (;;) {
        int64_t i = (int64_t)atomic_fetch_add(&b->next, 1);
        if (i >= b->n) break;
        int st;
        b->out[i] = orc_permanent_exact_square(b->mats + i * b->dim * b->dim, b->dim, &st);
    }
    return NULL;
}

double orc_permanent_batch(const double* mats, int64_t dim, int64_t n, int nThreads, double* out) {
    PermBatch b = {mats, dim, n, out, 0};
    if (nThreads < 1) nThreads = 1;
    pthread_t* th = (pthread_t*)malloc((size_t)nThreads * sizeof(pthread_t));
    double t0 = now_s();
    for (int t = 0; t < nThreads; t++) pthread_create(&th[t], NULL, perm_worker, &b);
    for (int t = 0; t < nThreads; t++) pthread_join(th[t], NULL);
    double dt = now_s() - t0;
    free(th);
    return dt;
}
