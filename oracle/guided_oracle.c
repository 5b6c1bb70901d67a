/*
 * oracle/guided_oracle.c -- TEST INFRASTRUCTURE (see oracle.h).  CPU restatement of the grid-guided projection matchers
 *   match::projection::match_frame_and_landmarks        src/stella_vslam/match/projection.cc:13-93   (mode 0)
 *   match::projection::match_current_and_last_frames    src/stella_vslam/match/projection.cc:95-207  (mode 1; the reprojection
 *                                                        itself, :122-157, is an input here)
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
    int n_matches = 0;
    for (int q = 0; q < Q; ++q) {
        match_out[q] = -1;
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
                    if (P->t_occupied[idx]) continue;
                    if (P->t_x_right && 0 < P->t_x_right[idx]) {
                        const float err = fabsf(P->q_x_right[q] - P->t_x_right[idx]);
                        if (margin < err) continue;
                    }
                    if (mode == 1 && check_orientation && fabsf(orc_angle_diff(P->q_angle[q], P->t_angle[idx])) > 30.0) continue;
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
        if (best <= thr) {
            if (mode == 0 && best_level == second_level && (float)best > lowe_ratio * (float)second) continue;
            match_out[q] = best_idx;
            P->t_occupied[best_idx] = 1; /* frm.add_landmark(...): the keypoint now carries an observed landmark */
            ++n_matches;
        }
    }
    free(cell_of); free(start); free(fill); free(items);
    return n_matches;
}
