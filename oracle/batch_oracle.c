/*
 * oracle/batch_oracle.c -- TEST INFRASTRUCTURE (see oracle.h).  Multi-threaded driver used ONLY as the timed CPU
 * baseline (bench.py cpu_baseline / --impl reference): n frames are extracted on a pthread pool (one frame per task,
 * like one stream per core), then every frame is brute-force matched against its predecessor.  The per-frame work is
 * the single-threaded restatement in orb_oracle.c / match_oracle.c (reference default: USE_OPENMP OFF).
 */
#include <malloc.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

typedef struct {
    const uint8_t* frames;
    int n, w, h, n_unique;
    const orc_orb_config_t* cfg;
    int cap;
    orc_keypoint_t* kps;
    uint8_t* descs;
    int* counts;
    float lowe;
    int check_ori;
    int* n_matches;
    volatile int next;
    int phase;
    pthread_mutex_t mu;
} batch_t;

static int take(batch_t* b) {
    pthread_mutex_lock(&b->mu);
    int i = b->next < b->n ? b->next++ : -1;
    pthread_mutex_unlock(&b->mu);
    return i;
}

static void* worker(void* arg) {
    batch_t* b = (batch_t*)arg;
    int i;
    float* ang1 = (float*)malloc(sizeof(float) * b->cap);
    float* ang2 = (float*)malloc(sizeof(float) * b->cap);
    int32_t* pairs = (int32_t*)malloc(sizeof(int32_t) * 2 * b->cap);
    while ((i = take(b)) >= 0) {
        if (b->phase == 0) {
            const uint8_t* img = b->frames + (size_t)(i % b->n_unique) * b->w * b->h;
            b->counts[i] = orc_orb_extract(img, b->w, b->h, b->w, NULL, 0, b->cfg, b->kps + (size_t)i * b->cap,
                                           b->descs + (size_t)i * b->cap * 32, b->cap, NULL, NULL, NULL);
        } else {
            const int j = (i + b->n - 1) % b->n;
            const int n1 = b->counts[i], n2 = b->counts[j];
            for (int k = 0; k < n1; ++k) ang1[k] = b->kps[(size_t)i * b->cap + k].angle;
            for (int k = 0; k < n2; ++k) ang2[k] = b->kps[(size_t)j * b->cap + k].angle;
            b->n_matches[i] = orc_brute_force_match(b->descs + (size_t)i * b->cap * 32, ang1, n1 < 0 ? 0 : n1, b->descs + (size_t)j * b->cap * 32,
                                                    ang2, NULL, n2 < 0 ? 0 : n2, b->lowe, b->check_ori, pairs);
        }
    }
    free(ang1);
    free(ang2);
    free(pairs);
    return NULL;
}

/* frames: n_unique tightly packed w*h images, cycled to make n frames.  counts/n_matches: [n] outputs.  Returns 0. */
int orc_frontend_batch(const uint8_t* frames, int n_unique, int n, int w, int h, const orc_orb_config_t* cfg, int cap, float lowe,
                       int check_ori, int n_threads, int* counts, int* n_matches) {
    /* keep the per-frame MB-sized scratch (pyramid, blur) on the per-thread heaps instead of mmap/munmap per call:
       with many threads the kernel's address-space lock otherwise serialises the workers */
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    mallopt(M_ARENA_MAX, 1024);
    batch_t b;
    memset(&b, 0, sizeof(b));
    b.frames = frames; b.n = n; b.w = w; b.h = h; b.n_unique = n_unique; b.cfg = cfg; b.cap = cap;
    b.kps = (orc_keypoint_t*)malloc(sizeof(orc_keypoint_t) * (size_t)n * cap);
    b.descs = (uint8_t*)malloc((size_t)n * cap * 32);
    b.counts = counts; b.n_matches = n_matches; b.lowe = lowe; b.check_ori = check_ori;
    pthread_mutex_init(&b.mu, NULL);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
    for (int phase = 0; phase < 2; ++phase) {
        b.phase = phase;
        b.next = 0;
        for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, worker, &b);
        for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    }
    free(th);
    free(b.kps);
    free(b.descs);
    pthread_mutex_destroy(&b.mu);
    return 0;
}
