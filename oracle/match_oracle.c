/*
 * oracle/match_oracle.c -- TEST INFRASTRUCTURE (see oracle.h).  CPU restatement of the 256-bit Hamming matchers.
 * Pinned by the reference's own known-answer tests (test/stella_vslam/match/base.cc:11-57) for the distance, and since round 2 by
 * cv2.BFMatcher(NORM_HAMMING) for the distance matrix and the best / second-best pair that brute_force_match consumes
 * (tests/golden/match_bf_cv2.npz, tests/test_natural_golden.py).  The greedy part of brute_force_match (taken set, ratio test,
 * orientation histogram, output order; robust.cc:253-325) has no reference test: "parity unpinned" beyond the restatement and its
 * literal Python cross-check.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* match::compute_descriptor_distance_32 (match/base.h:20-41): SWAR popcount over 8 x u32. */
unsigned orc_hamming_32(const uint8_t* a, const uint8_t* b) {
    unsigned dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t x, y;
        memcpy(&x, a + 4 * i, 4);
        memcpy(&y, b + 4 * i, 4);
        uint32_t v = x ^ y;
        v -= ((v >> 1) & 0x55555555U);
        v = (v & 0x33333333U) + ((v >> 2) & 0x33333333U);
        dist += (((v + (v >> 4)) & 0x0F0F0F0FU) * 0x01010101U) >> 24;
    }
    return dist;
}

/* match::compute_descriptor_distance_64 (match/base.h:44-65): SWAR popcount over 4 x u64. */
unsigned orc_hamming_64(const uint8_t* a, const uint8_t* b) {
    unsigned dist = 0;
    for (int i = 0; i < 4; ++i) {
        uint64_t x, y;
        memcpy(&x, a + 8 * i, 8);
        memcpy(&y, b + 8 * i, 8);
        uint64_t v = x ^ y;
        v -= (v >> 1) & 0x5555555555555555ULL;
        v = (v & 0x3333333333333333ULL) + ((v >> 2) & 0x3333333333333333ULL);
        dist += (unsigned)((((v + (v >> 4)) & 0x0F0F0F0F0F0F0F0FULL) * 0x0101010101010101ULL) >> 56);
    }
    return dist;
}

/* util::angle::diff (util/angle.cc:7-16): float subtraction, comparisons against double literals. */
float orc_angle_diff(float a1, float a2) {
    float ret = a1 - a2;
    if (ret <= -180.0) ret += 360.0;
    if (ret > 180.0) ret -= 360.0;
    return ret;
}

/* match::robust::brute_force_match (match/robust.cc:232-328). */
int orc_brute_force_match(const uint8_t* desc1, const float* angle1, int n1, const uint8_t* desc2, const float* angle2,
                          const uint8_t* valid2, int n2, float lowe_ratio, int check_orientation, int32_t* pairs_out) {
    int* matched_2_in_1 = (int*)malloc(sizeof(int) * (n1 ? n1 : 1));
    uint8_t* taken_1 = (uint8_t*)calloc(n1 ? n1 : 1, 1);
    for (int i = 0; i < n1; ++i) matched_2_in_1[i] = -1;
    for (int idx_2 = 0; idx_2 < n2; ++idx_2) {
        if (valid2 && !valid2[idx_2]) continue; /* no landmark / will_be_erased (robust.cc:255-262) */
        const uint8_t* d2 = desc2 + (size_t)idx_2 * 32;
        unsigned best = 256, second = 256;
        int best_idx_1 = -1;
        for (int idx_1 = 0; idx_1 < n1; ++idx_1) {
            if (taken_1[idx_1]) continue;
            if (check_orientation && fabsf(orc_angle_diff(angle1[idx_1], angle2[idx_2])) > 30.0) continue;
            const unsigned hd = orc_hamming_32(d2, desc1 + (size_t)idx_1 * 32);
            if (hd < best) {
                second = best;
                best = hd;
                best_idx_1 = idx_1;
            } else if (hd < second) {
                second = hd;
            }
        }
        if (50 < best) continue;
        if (best_idx_1 < 0) continue;
        if (lowe_ratio * second < (float)best) continue;
        matched_2_in_1[best_idx_1] = idx_2;
        taken_1[best_idx_1] = 1;
    }
    int n = 0;
    for (int idx_1 = 0; idx_1 < n1; ++idx_1) {
        if (matched_2_in_1[idx_1] < 0) continue;
        pairs_out[2 * n] = idx_1;
        pairs_out[2 * n + 1] = matched_2_in_1[idx_1];
        ++n;
    }
    free(matched_2_in_1);
    free(taken_1);
    return n;
}

/* data::landmark::compute_descriptor (src/stella_vslam/data/landmark.cc:199-256): among the descriptors of a landmark's observations,
 * the one whose median Hamming distance to all of them (itself included; element [0.5 (n-1)] of the sorted row) is smallest, first
 * index on ties.  descs: n x 32.  Returns best_idx. */
static int cmp_uint(const void* a, const void* b) {
    const unsigned x = *(const unsigned*)a, y = *(const unsigned*)b;
    return x < y ? -1 : (x > y);
}
int orc_landmark_descriptor(const uint8_t* descs, int n) {
    if (n <= 0) return -1;
    unsigned* d = (unsigned*)malloc(sizeof(unsigned) * (size_t)n * n);
    unsigned* row = (unsigned*)malloc(sizeof(unsigned) * (size_t)n);
    for (int i = 0; i < n; ++i) {
        d[(size_t)i * n + i] = 0;
        for (int j = i + 1; j < n; ++j) d[(size_t)i * n + j] = d[(size_t)j * n + i] = orc_hamming_32(descs + 32 * (size_t)i, descs + 32 * (size_t)j);
    }
    unsigned best = 256;
    int best_idx = 0;
    for (int i = 0; i < n; ++i) {
        memcpy(row, d + (size_t)i * n, sizeof(unsigned) * (size_t)n);
        qsort(row, (size_t)n, sizeof(unsigned), cmp_uint);
        const unsigned med = row[(unsigned)(0.5 * (n - 1))];
        if (med < best) {
            best = med;
            best_idx = i;
        }
    }
    free(d); free(row);
    return best_idx;
}
