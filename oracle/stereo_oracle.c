/*
 * oracle/stereo_oracle.c -- TEST INFRASTRUCTURE (see oracle.h).  CPU restatement of match::stereo
 *   stereo::compute                                   src/stella_vslam/match/stereo.cc:20-114
 *   stereo::get_right_keypoint_indices_in_each_row    stereo.cc:116-142
 *   stereo::find_closest_keypoints_in_stereo          stereo.cc:144-178
 *   stereo::compute_subpixel_disparity                stereo.cc:180-251
 * (cv::norm(NORM_L1) on the two centre-subtracted CV_32F patches sums |a - b| of integer-valued floats in double: exact, so the
 * correlation is an integer).  The reference has no test for this matcher: parity unpinned beyond this restatement.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

static inline int floor_f(float v) {
    int i = (int)v;
    return i - (i > v);
}
static inline int ceil_f(float v) {
    int i = (int)v;
    return i + (i < v);
}

typedef struct {
    int corr, idx;
} corr_item_t;
static int cmp_corr(const void* a, const void* b) {
    const corr_item_t *x = (const corr_item_t*)a, *y = (const corr_item_t*)b;
    if (x->corr != y->corr) return x->corr < y->corr ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

int orc_stereo_compute(const uint8_t* const* pyr_left, const uint8_t* const* pyr_right, const int* widths, const int* heights,
                       const orc_keypoint_t* kl, const uint8_t* dl, int n_left, const orc_keypoint_t* kr, const uint8_t* dr, int n_right,
                       const float* scale_factors, const float* inv_scale_factors, float focal_x_baseline, float true_baseline,
                       float* stereo_x_right, float* depths) {
    const unsigned thr = (100u + 50u) / 2u; /* stereo.h:99 */
    const float min_disp = 0.0f, max_disp = focal_x_baseline / true_baseline;
    const int rows = heights[0];
    /* get_right_keypoint_indices_in_each_row(2.0): CSR of right keypoint indices per image row, ascending index inside a row */
    int* start = (int*)calloc((size_t)rows + 1, sizeof(int));
    int* lo = (int*)malloc(sizeof(int) * (size_t)(n_right > 0 ? n_right : 1));
    int* hi = (int*)malloc(sizeof(int) * (size_t)(n_right > 0 ? n_right : 1));
    for (int j = 0; j < n_right; ++j) {
        const float r = 2.0f * scale_factors[kr[j].octave];
        hi[j] = ceil_f(kr[j].y + r);
        lo[j] = floor_f(kr[j].y - r);
        for (int row = lo[j]; row <= hi[j]; ++row)
            if (0 <= row && row < rows) start[row + 1]++; /* the reference's .at() would throw outside; the extractor never puts keypoints there */
    }
    for (int r = 0; r < rows; ++r) start[r + 1] += start[r];
    int* items = (int*)malloc(sizeof(int) * (size_t)(start[rows] > 0 ? start[rows] : 1));
    int* fill = (int*)malloc(sizeof(int) * (size_t)(rows > 0 ? rows : 1));
    memcpy(fill, start, sizeof(int) * (size_t)rows);
    for (int j = 0; j < n_right; ++j)
        for (int row = lo[j]; row <= hi[j]; ++row)
            if (0 <= row && row < rows) items[fill[row]++] = j;

    corr_item_t* corr = (corr_item_t*)malloc(sizeof(corr_item_t) * (size_t)(n_left > 0 ? n_left : 1));
    int n_corr = 0;
    for (int i = 0; i < n_left; ++i) {
        stereo_x_right[i] = -1.0f;
        depths[i] = -1.0f;
    }
    for (int i = 0; i < n_left; ++i) {
        const int level = kl[i].octave;
        const float x_left = kl[i].x, y_left = kl[i].y;
        const size_t row = (size_t)y_left;
        if (row >= (size_t)rows || start[row] == start[row + 1]) continue;
        const float min_x_right = x_left - max_disp, max_x_right = x_left - min_disp;
        if (max_x_right < 0) continue;
        unsigned best = thr, best_j = 0;
        for (int k = start[row]; k < start[row + 1]; ++k) {
            const int j = items[k];
            if (kr[j].octave < level - 1 || kr[j].octave > level + 1) continue;
            if (kr[j].x < min_x_right || max_x_right < kr[j].x) continue;
            const unsigned d = orc_hamming_32(dl + (size_t)i * 32, dr + (size_t)j * 32);
            if (d < best) {
                best = d;
                best_j = (unsigned)j;
            }
        }
        if (thr <= best) continue;
        /* compute_subpixel_disparity */
        const float x_right = kr[best_j].x, inv_sf = inv_scale_factors[level];
        const int sxl = (int)lrintf(x_left * inv_sf), syl = (int)lrintf(y_left * inv_sf), sxr = (int)lrintf(x_right * inv_sf);
        const int win = 5, slide = 5;
        const int ini_x = sxr - slide - win, end_x = sxr + slide + win;
        const int W = widths[level];
        if (ini_x < 0 || W <= end_x) continue;
        const uint8_t *L = pyr_left[level], *R = pyr_right[level];
        const int lc = L[(size_t)syl * W + sxl];
        int correlations[11], best_corr = 0x7FFFFFFF, best_offset = 0;
        for (int off = -slide; off <= slide; ++off) {
            const int rc = R[(size_t)syl * W + sxr + off];
            int s = 0;
            for (int dy = -win; dy <= win; ++dy)
                for (int dx = -win; dx <= win; ++dx)
                    s += abs((L[(size_t)(syl + dy) * W + sxl + dx] - lc) - (R[(size_t)(syl + dy) * W + sxr + off + dx] - rc));
            if (s < best_corr) {
                best_corr = s;
                best_offset = off;
            }
            correlations[slide + off] = s;
        }
        if (best_offset == -slide || best_offset == slide) continue;
        const float c1 = (float)correlations[slide + best_offset - 1], c2 = (float)correlations[slide + best_offset],
                    c3 = (float)correlations[slide + best_offset + 1];
        const float x_delta = (float)((c1 - c3) / (2.0 * (c1 + c3) - 4.0 * c2));
        if (x_delta < -1.0 || 1.0 < x_delta) continue;
        float best_x_right = scale_factors[level] * ((float)(sxr + best_offset) + x_delta);
        float best_disp = x_left - best_x_right;
        if (best_disp < min_disp || max_disp <= best_disp) continue;
        if (best_disp <= 0.0f) {
            best_disp = 0.01f;
            best_x_right = x_left - best_disp;
        }
        depths[i] = focal_x_baseline / best_disp;
        stereo_x_right[i] = best_x_right;
        corr[n_corr].corr = best_corr;
        corr[n_corr].idx = i;
        ++n_corr;
    }
    /* stereo.cc:96-113: reject matches whose correlation exceeds twice the median */
    qsort(corr, (size_t)n_corr, sizeof(corr_item_t), cmp_corr);
    const int median_i = n_corr / 2;
    const float median = n_corr == 0 ? 0.0f : (float)corr[median_i].corr;
    const float corr_thr = (float)(2.0 * median);
    int kept = n_corr;
    for (int k = median_i; k < n_corr; ++k)
        if (corr_thr < (float)corr[k].corr) {
            stereo_x_right[corr[k].idx] = -1;
            depths[corr[k].idx] = -1;
            --kept;
        }
    free(start); free(lo); free(hi); free(items); free(fill); free(corr);
    return kept;
}
