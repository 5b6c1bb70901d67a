/*
 * oracle/pairs_oracle.c -- TEST INFRASTRUCTURE (see oracle.h).  CPU restatement of the all-pairs matchers that carry greedy state
 *   variant 0 (BOW):            match::bow_tree::match_frame_and_keyframe   src/stella_vslam/match/bow_tree.cc:169-256
 *                               match::bow_tree::match_keyframes            bow_tree.cc:258-366
 *   variant 1 (TRIANGULATION):  match::robust::match_for_triangulation      src/stella_vslam/match/robust.cc:14-146
 *                               match::bow_tree::match_for_triangulation    bow_tree.cc:11-167
 *   check_epipolar_constraint                                               match/base.h:67-79
 * When node ids are given, rows are visited exactly like the merge-join over the two bow_feature_vectors (ascending node id,
 * then ascending keypoint index inside the node); without them every row sees every candidate (robust::).
 * The reference has no test for these matchers: parity unpinned beyond this restatement.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

int orc_check_epipolar_constraint(const double* b1, const double* b2, const double* E, float residual_rad_thr, float scale_factor) {
    /* epiplane_in_1 = E_12 * bearing_2 (3x3 row-major times 3) */
    const double e0 = E[0] * b2[0] + E[1] * b2[1] + E[2] * b2[2];
    const double e1 = E[3] * b2[0] + E[4] * b2[1] + E[5] * b2[2];
    const double e2 = E[6] * b2[0] + E[7] * b2[1] + E[8] * b2[2];
    const double dot = e0 * b1[0] + e1 * b1[1] + e2 * b1[2];
    const double norm = sqrt(e0 * e0 + e1 * e1 + e2 * e2);
    double c = dot / norm;
    c = fmax(-1.0, c);
    c = fmin(1.0, c);
    const double residual_rad = fabs(M_PI / 2.0 - acos(c));
    return residual_rad < residual_rad_thr * scale_factor; /* float product, compared in double */
}

typedef struct {
    int node, idx;
} node_item_t;
static int cmp_node_item(const void* a, const void* b) {
    const node_item_t *x = (const node_item_t*)a, *y = (const node_item_t*)b;
    if (x->node != y->node) return x->node < y->node ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

int orc_match_pairs(const orc_pairs_t* P, int variant, float lowe_ratio, int check_orientation, int32_t* match_out) {
    const int N1 = P->n1, N2 = P->n2;
    uint8_t* taken = (uint8_t*)calloc((size_t)(N2 > 0 ? N2 : 1), 1);
    node_item_t* rows = (node_item_t*)malloc(sizeof(node_item_t) * (size_t)(N1 > 0 ? N1 : 1));
    node_item_t* cands = (node_item_t*)malloc(sizeof(node_item_t) * (size_t)(N2 > 0 ? N2 : 1));
    for (int i = 0; i < N1; ++i) {
        rows[i].node = P->node1 ? P->node1[i] : 0;
        rows[i].idx = i;
        match_out[i] = -1;
    }
    for (int j = 0; j < N2; ++j) {
        cands[j].node = P->node2 ? P->node2[j] : 0;
        cands[j].idx = j;
    }
    qsort(rows, (size_t)N1, sizeof(node_item_t), cmp_node_item);
    qsort(cands, (size_t)N2, sizeof(node_item_t), cmp_node_item);
    int n_matches = 0, c_lo = 0;
    for (int r = 0; r < N1; ++r) {
        const int i = rows[r].idx, node = rows[r].node;
        while (c_lo < N2 && cands[c_lo].node < node) ++c_lo; /* lower_bound of the merge-join */
        if (P->valid1 && !P->valid1[i]) continue;
        const uint8_t* d1 = P->desc1 + (size_t)i * 32;
        unsigned best = variant == 1 ? 50u : 256u, second = 256u;
        int best_idx = -1;
        for (int c = c_lo; c < N2 && cands[c].node == node; ++c) {
            const int j = cands[c].idx;
            if (P->valid2 && !P->valid2[j]) continue;
            if (taken[j]) continue;
            if (check_orientation && fabsf(orc_angle_diff(P->angle1[i], P->angle2[j])) > 30.0) continue;
            const unsigned d = orc_hamming_32(d1, P->desc2 + (size_t)j * 32);
            if (variant == 1) {
                if (50u < d || best < d) continue;
                const int stereo1 = P->stereo1 && P->stereo1[i], stereo2 = P->stereo2 && P->stereo2[j];
                const double* b2 = P->bearing2 + (size_t)j * 3;
                if (P->valid_epiplane && !stereo1 && !stereo2) {
                    const double cos_dist = P->epiplane_in_2[0] * b2[0] + P->epiplane_in_2[1] * b2[1] + P->epiplane_in_2[2] * b2[2];
                    if (0.99862953475 < cos_dist) continue;
                }
                if (!orc_check_epipolar_constraint(P->bearing1 + (size_t)i * 3, b2, P->E_12, P->residual_rad_thr, P->scale1[i])) continue;
            }
            if (d < best) {
                second = best;
                best = d;
                best_idx = j;
            } else if (d < second) {
                second = d;
            }
        }
        if (variant == 1) {
            if (best_idx < 0) continue;
        } else if (50u < best) {
            continue;
        }
        if (lowe_ratio * (float)second < (float)best) continue;
        taken[best_idx] = 1;
        match_out[i] = best_idx;
        ++n_matches;
    }
    free(taken); free(rows); free(cands);
    return n_matches;
}
