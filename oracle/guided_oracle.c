/*
 * oracle/guided_oracle.c -- TEST INFRASTRUCTURE (see oracle.h).  CPU restatement of the grid-guided projection matchers
 *   match::projection::match_frame_and_landmarks        src/stella_vslam/match/projection.cc:13-93   (mode 0)
 *   match::projection::match_current_and_last_frames    src/stella_vslam/match/projection.cc:95-207  (mode 1; the reprojection
 *                                                        itself, :122-157, is an input here)
 *   match::projection::match_frame_and_keyframe         projection.cc:217-319   (mode 1, thr = hamm_dist_thr, no stereo gate)
 *   match::projection::match_by_Sim3_transform          projection.cc:321-416   (mode 1, thr = 50, no orientation)
 *   match::projection::match_keyframes_mutually         projection.cc:418-630   (mode 2 twice + orc_cross_check)
 *   match::fuse::detect_duplication                     match/fuse.cc:12-154    (mode 3)
 *   match::area::match_in_consistent_area               match/area.cc:8-98      (mode 4)
 * with data::assign_keypoints_to_grid / get_keypoints_in_cell (data/common.cc:83-190, data/common.h:60-68).
 * The reference has no test for these matchers: parity unpinned beyond this restatement.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

static inline int floor_d(double v) {
    int i = (int)v;
    return i - (v < (double)i);
}
static inline int ceil_d(double v) {
    int i = (int)v;
    return i + (v > (double)i);
}

int orc_match_guided(const orc_guided_t* P, int mode, unsigned thr, float lowe_ratio, int check_orientation, int32_t* match_out) {
    const int N = P->n_train, Q = P->n_queries, GC = P->grid_cols, GR = P->grid_rows;
    const double inv_w = (double)GC / (P->max_x - P->min_x), inv_h = (double)GR / (P->max_y - P->min_y);
    /* assign_keypoints_to_grid: cells[x][y] hold keypoint indices in ascending order */
    int* cell_of = (int*)malloc(sizeof(int) * (N ? N : 1));
    int* start = (int*)calloc((size_t)GC * GR + 1, sizeof(int));
    for (int i = 0; i < N; ++i) {
        const int cx = floor_d((P->t_x[i] - P->min_x) * inv_w), cy = floor_d((P->t_y[i] - P->min_y) * inv_h);
        cell_of[i] = (0 <= cx && cx < GC && 0 <= cy && cy < GR) ? cx * GR + cy : -1;
        if (cell_of[i] >= 0) start[cell_of[i] + 1]++;
    }
    for (int c = 0; c < GC * GR; ++c) start[c + 1] += start[c];
    int* fill = (int*)malloc(sizeof(int) * (size_t)GC * GR);
    memcpy(fill, start, sizeof(int) * (size_t)GC * GR);
    int* items = (int*)malloc(sizeof(int) * (N ? N : 1));
    for (int i = 0; i < N; ++i)
        if (cell_of[i] >= 0) items[fill[cell_of[i]]++] = i;
    /* per-keypoint state: 0 = occupied, 0xFFFF = free; in mode 4 the Hamming distance of the match it currently holds (256 = none) */
    uint16_t* state = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)(N > 0 ? N : 1));
    int* owner = (int*)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
    for (int i = 0; i < N; ++i) {
        state[i] = mode == 4 ? 256 : ((P->t_occupied && P->t_occupied[i]) ? 0 : 0xFFFF);
        owner[i] = -1;
    }
    int n_matches = 0;
    for (int q = 0; q < Q; ++q) match_out[q] = -1;
    for (int q = 0; q < Q; ++q) {
        if (P->q_valid && !P->q_valid[q]) continue;
        const float ref_x = P->q_x[q], ref_y = P->q_y[q], margin = P->q_margin[q];
        const int min_level = P->q_min_level[q], max_level = P->q_max_level[q];
        /* get_keypoints_in_cell (data/common.cc:127-190) */
        int min_cx = floor_d((double)(ref_x - P->min_x - margin) * inv_w);
        if (min_cx < 0) min_cx = 0;
        if (GC <= min_cx) continue;
        int max_cx = ceil_d((double)(ref_x - P->min_x + margin) * inv_w);
        if (max_cx > GC - 1) max_cx = GC - 1;
        if (max_cx < 0) continue;
        int min_cy = floor_d((double)(ref_y - P->min_y - margin) * inv_h);
        if (min_cy < 0) min_cy = 0;
        if (GR <= min_cy) continue;
        int max_cy = ceil_d((double)(ref_y - P->min_y + margin) * inv_h);
        if (max_cy > GR - 1) max_cy = GR - 1;
        if (max_cy < 0) continue;
        unsigned best = 256, second = 256;
        int best_level = -1, second_level = -1, best_idx = -1;
        for (int cx = min_cx; cx <= max_cx; ++cx)
            for (int cy = min_cy; cy <= max_cy; ++cy)
                for (int k = start[cx * GR + cy]; k < start[cx * GR + cy + 1]; ++k) {
                    const int idx = items[k];
                    const int oct = P->t_octave[idx];
                    if (0 <= min_level && oct < min_level) continue;
                    if (0 <= max_level && max_level < oct) continue;
                    const float dx = P->t_x[idx] - ref_x, dy = P->t_y[idx] - ref_y;
                    if (!(fabsf(dx) < margin && fabsf(dy) < margin)) continue;
                    /* matcher loop */
                    if (mode == 4) {
                        if (check_orientation && fabsf(orc_angle_diff(P->q_angle[q], P->t_angle[idx])) > 30.0) continue;
                        const unsigned d4 = orc_hamming_32(P->q_desc + (size_t)q * 32, P->t_desc + (size_t)idx * 32);
                        if (state[idx] <= d4) continue; /* area.cc:49-51: the already-matched point is closer */
                        if (d4 < best) {
                            second = best;
                            best = d4;
                            best_idx = idx;
                        } else if (d4 < second) {
                            second = d4;
                        }
                        continue;
                    }
                    if (state[idx] == 0) continue;
                    if (mode <= 1 && P->t_x_right && 0 < P->t_x_right[idx]) {
                        const float err = fabsf(P->q_x_right[q] - P->t_x_right[idx]);
                        if (margin < err) continue;
                    }
                    if (mode == 1 && check_orientation && fabsf(orc_angle_diff(P->q_angle[q], P->t_angle[idx])) > 30.0) continue;
                    if (mode == 3 && P->do_reprojection_matching) { /* fuse.cc:93-120 */
                        const double e_x = P->q_reproj[2 * q] - P->t_x[idx], e_y = P->q_reproj[2 * q + 1] - P->t_y[idx];
                        const float inv_sigma_sq = P->inv_level_sigma_sq[oct];
                        if (P->t_x_right && P->t_x_right[idx] >= 0) {
                            const float e_xr = P->q_x_right[q] - P->t_x_right[idx];
                            const double err_sq = e_x * e_x + e_y * e_y + e_xr * e_xr;
                            const float chi_sq_3D = 7.81473;
                            if (chi_sq_3D < err_sq * inv_sigma_sq) continue;
                        } else {
                            const double err_sq = e_x * e_x + e_y * e_y;
                            const float chi_sq_2D = 5.99146;
                            if (chi_sq_2D < err_sq * inv_sigma_sq) continue;
                        }
                    }
                    const unsigned d = orc_hamming_32(P->q_desc + (size_t)q * 32, P->t_desc + (size_t)idx * 32);
                    if (d < best) {
                        second = best;
                        second_level = best_level;
                        best = d;
                        best_level = oct;
                        best_idx = idx;
                    } else if (mode == 0 && d < second) {
                        second_level = oct;
                        second = d;
                    }
                }
        if (best_idx < 0) continue; /* indices empty or everything gated: best stays 256 > thr */
        if (best > thr) continue;
        if (mode == 0 && best_level == second_level && (float)best > lowe_ratio * (float)second) continue;
        if (mode == 4) {
            if ((float)second * lowe_ratio < (float)best) continue; /* area.cc:66-68 */
            const int prev = owner[best_idx];                       /* area.cc:75-81: steal the keypoint from its previous owner */
            if (0 <= prev) {
                match_out[prev] = -1;
                --n_matches;
            }
            owner[best_idx] = q;
            state[best_idx] = (uint16_t)best;
        } else if (mode != 2) {
            /* frm.add_landmark(...) / matched_lms.at(idx) = lm / already_matched_idx.insert(idx).  The later test is
             * `lm && lm->has_observation()` (projection.cc:50-53, 163-166): a landmark without observations does not close the keypoint */
            if (!P->q_has_observation || P->q_has_observation[q]) state[best_idx] = 0;
        }
        match_out[q] = best_idx;
        ++n_matches;
    }
    if (P->t_occupied && mode != 2 && mode != 4)
        for (int i = 0; i < N; ++i) P->t_occupied[i] = state[i] == 0;
    free(state); free(owner);
    free(cell_of); free(start); free(fill); free(items);
    return n_matches;
}

/* the closing loop of match_keyframes_mutually (projection.cc:614-627): keep i -> j only when j -> i */
int orc_cross_check(const int32_t* idx2_in_1, int n1, const int32_t* idx1_in_2, int n2, int32_t* mutual_out) {
    int n = 0;
    for (int i = 0; i < n1; ++i) {
        mutual_out[i] = -1;
        const int j = idx2_in_1[i];
        if (j < 0 || j >= n2) continue;
        if (idx1_in_2[j] == i) {
            mutual_out[i] = j;
            ++n;
        }
    }
    return n;
}
