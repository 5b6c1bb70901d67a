/*
 * oracle/orb_oracle.c -- TEST INFRASTRUCTURE (see oracle.h).  CPU restatement of feature::orb_extractor.
 *
 * Follows src/stella_vslam/feature/orb_extractor.cc, orb_impl.cc, orb_params.cc, util/trigonometric.h and
 * restates the four OpenCV primitives the reference calls (not vendored in /root/reference; Docker pin 4.7.0,
 * validated here against the cv2 4.13.0 wheel -- see tests/golden/make_golden.py and tests/test_oracle_cv2.py):
 *   cv::resize(INTER_LINEAR, 8UC1)           orb_extractor.cc:160
 *   cv::FAST(thr, nonmax=true) (TYPE_9_16)   orb_extractor.cc:228-235
 *   cv::GaussianBlur(7x7, sigma 2, REFLECT_101) orb_extractor.cc:103
 *   cv::fastAtan2                            orb_impl.cc:90
 * Compile with -ffp-contract=off: the reference's fp32 arithmetic has no fused multiply-add.
 */
#include "oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

static const int8_t k_pattern[1024] = {
#include "../stella_vslam_b200/csrc/orb_pattern.inc"
};

/* cvRound on x86-64 = cvtss2si / cvtsd2si under the default rounding mode: round half to even. */
static inline int round_half_even_f(float v) { return (int)lrintf(v); }
static inline int floor_f(float v) {
    int i = (int)v;
    return i - (v < (float)i);
}

/* ------------------------------------------------------------------------------------------------
 * orb_params::calc_* (orb_params.cc:37-71): float recurrences, not pow().
 * ---------------------------------------------------------------------------------------------- */
void orc_scale_factors(float scale_factor, int n, float* sf, float* inv_sf, float* sigma_sq, float* inv_sigma_sq) {
    float s = 1.0f, is = 1.0f;
    for (int l = 0; l < n; ++l) {
        if (l > 0) {
            s = scale_factor * s;
            is = (1.0f / scale_factor) * is;
        }
        if (sf) sf[l] = s;
        if (inv_sf) inv_sf[l] = is;
        if (sigma_sq) sigma_sq[l] = (l == 0) ? 1.0f : s * s;
        if (inv_sigma_sq) inv_sigma_sq[l] = (l == 0) ? 1.0f : 1.0f / (s * s);
    }
}

/* orb_extractor.cc:157-158: sizes come from the ORIGINAL image and the float scale factor widened to double. */
void orc_level_size(int w, int h, float sf, int* lw, int* lh) {
    const double scale = sf;
    *lw = (int)round(w * 1.0 / scale);
    *lh = (int)round(h * 1.0 / scale);
}

/* ------------------------------------------------------------------------------------------------
 * cv::resize INTER_LINEAR 8UC1 (OpenCV imgproc/resize.cpp: resizeGeneric_ + HResizeLinear/VResizeLinear,
 * INTER_RESIZE_COEF_BITS = 11).  Horizontal taps clamp fx at the borders, vertical rows are clipped.
 * ---------------------------------------------------------------------------------------------- */
static inline short sat_short_round(float v) {
    long r = lrintf(v);
    if (r > 32767) r = 32767;
    if (r < -32768) r = -32768;
    return (short)r;
}

void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride) {
    const double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
    const double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
    int* xofs = (int*)malloc(sizeof(int) * dw);
    short* alpha = (short*)malloc(sizeof(short) * 2 * dw);
    int* row0 = (int*)malloc(sizeof(int) * dw);
    int* row1 = (int*)malloc(sizeof(int) * dw);
    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = floor_f(fx);
        fx -= sx;
        if (sx < 0) {
            fx = 0;
            sx = 0;
        }
        if (sx >= sw - 1) {
            fx = 0;
            sx = sw - 1;
        }
        xofs[dx] = sx;
        alpha[2 * dx] = sat_short_round((1.f - fx) * 2048.f);
        alpha[2 * dx + 1] = sat_short_round(fx * 2048.f);
    }
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = floor_f(fy);
        fy -= sy;
        const short b0 = sat_short_round((1.f - fy) * 2048.f), b1 = sat_short_round(fy * 2048.f);
        int sy0 = sy < 0 ? 0 : (sy < sh ? sy : sh - 1);
        int sy1 = sy + 1 < 0 ? 0 : (sy + 1 < sh ? sy + 1 : sh - 1);
        const uint8_t* s0 = src + (size_t)sy0 * sstride;
        const uint8_t* s1 = src + (size_t)sy1 * sstride;
        for (int dx = 0; dx < dw; ++dx) {
            const int sx = xofs[dx];
            const int sx1 = sx + 1 < sw ? sx + 1 : sx; /* weight is 0 there */
            row0[dx] = s0[sx] * alpha[2 * dx] + s0[sx1] * alpha[2 * dx + 1];
            row1[dx] = s1[sx] * alpha[2 * dx] + s1[sx1] * alpha[2 * dx + 1];
        }
        uint8_t* d = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; ++dx) {
            int v = (((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2;
            d[dx] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
    free(xofs);
    free(alpha);
    free(row0);
    free(row1);
}

/* ------------------------------------------------------------------------------------------------
 * cv::FAST TYPE_9_16 with non-max suppression on a (sub-)image treated as a whole image
 * (OpenCV features2d/fast.cpp FAST_t<16> + fast_score.cpp cornerScore<16>).
 * Candidates x in [3,w-4], y in [3,h-4].  m = max over the 16 arcs of 9 contiguous circle pixels of
 * min(v-p) (darker arc) / min(p-v) (brighter arc); corner at thr <=> m > thr; score = m-1.  NMS keeps a corner
 * iff its score is strictly greater than the 8 neighbours' scores (non-corners and pixels outside the candidate
 * region score 0).  Output is row-major (y, then x).
 * ---------------------------------------------------------------------------------------------- */
static const int k_circle_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int k_circle_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

static inline int fast_m_value(const uint8_t* p, const int* off, int thr) {
    const int v = p[0];
    /* any 9-arc contains circle pixel 0 or 8, and 4 or 12: cheap rejection first */
    const int d0 = v - p[off[0]], d8 = v - p[off[8]];
    if (!((d0 > thr) | (d0 < -thr) | (d8 > thr) | (d8 < -thr))) return 0;
    const int d4 = v - p[off[4]], d12 = v - p[off[12]];
    if (!((d4 > thr) | (d4 < -thr) | (d12 > thr) | (d12 < -thr))) return 0;
    int d[25];
    for (int k = 0; k < 16; ++k) d[k] = v - p[off[k]];
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    int best = 0;
    for (int s = 0; s < 16; ++s) {
        int lo = d[s], hi = d[s];
        for (int k = 1; k < 9; ++k) {
            if (d[s + k] < lo) lo = d[s + k];
            if (d[s + k] > hi) hi = d[s + k];
        }
        if (lo > best) best = lo;   /* all nine darker than v by >= lo */
        if (-hi > best) best = -hi; /* all nine brighter than v by >= -hi */
    }
    return best;
}

int orc_fast9_16_nms(const uint8_t* img, int stride, int w, int h, int thr, int16_t* xs, int16_t* ys, uint8_t* scores, int cap) {
    if (w < 7 || h < 7) return 0;
    if (thr < 0) thr = 0;
    if (thr > 255) thr = 255;
    int off[16];
    for (int k = 0; k < 16; ++k) off[k] = k_circle_dy[k] * stride + k_circle_dx[k];
    uint8_t* sc = (uint8_t*)calloc((size_t)w * h, 1);
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            const int m = fast_m_value(img + (size_t)y * stride + x, off, thr);
            if (m > thr) sc[y * w + x] = (uint8_t)(m - 1);
        }
    int n = 0;
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            const int s = sc[y * w + x];
            if (s == 0 && thr > 0) continue;
            if (s == 0) continue; /* thr 0: m>0 => score m-1 may be 0; cv keeps score>all neighbours strictly, 0 never wins */
            const uint8_t* r = sc + y * w + x;
            if (s > r[-1] && s > r[1] && s > r[-w - 1] && s > r[-w] && s > r[-w + 1] && s > r[w - 1] && s > r[w] && s > r[w + 1]) {
                if (n < cap) {
                    xs[n] = (int16_t)x;
                    ys[n] = (int16_t)y;
                    scores[n] = (uint8_t)s;
                }
                ++n;
            }
        }
    free(sc);
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * cv::GaussianBlur(7x7, sigma=2, BORDER_REFLECT_101) on CV_8U: OpenCV >= 3.4.2 fixed-point path
 * (imgproc/smooth.dispatch.cpp + fixedpoint.inl.hpp): Q8.8 taps {18,34,48,56,48,34,18}/256,
 * horizontal pass in 16 bit, vertical pass in Q16.16, round to nearest, saturate.
 * ---------------------------------------------------------------------------------------------- */
static inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * n - 2 - i;
    }
    return i;
}

void orc_gaussian7_s2_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
    static const int tap[7] = {18, 34, 48, 56, 48, 34, 18};
    uint16_t* hbuf = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)w * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t* s = src + (size_t)y * sstride;
        uint16_t* hb = hbuf + (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            unsigned acc = 0;
            if (x >= 3 && x < w - 3) {
                for (int k = 0; k < 7; ++k) acc += tap[k] * s[x + k - 3];
            } else {
                for (int k = 0; k < 7; ++k) acc += tap[k] * s[reflect101(x + k - 3, w)];
            }
            hb[x] = (uint16_t)acc;
        }
    }
    for (int y = 0; y < h; ++y) {
        const uint16_t* r[7];
        for (int k = 0; k < 7; ++k) r[k] = hbuf + (size_t)reflect101(y + k - 3, h) * w;
        uint8_t* d = dst + (size_t)y * dstride;
        for (int x = 0; x < w; ++x) {
            uint32_t acc = 0;
            for (int k = 0; k < 7; ++k) acc += (uint32_t)tap[k] * r[k][x];
            acc = (acc + (1u << 15)) >> 16;
            d[x] = (uint8_t)(acc > 255 ? 255 : acc);
        }
    }
    free(hbuf);
}

/* ------------------------------------------------------------------------------------------------
 * cv::fastAtan2 scalar path (OpenCV core/mathfuncs_core.simd.hpp atan_f32): degrees in [0,360).
 * ---------------------------------------------------------------------------------------------- */
float orc_fast_atan2(float y, float x) {
    const float rad2deg = (float)(180 / 3.1415926535897932384626433832795);
    const float p1 = 0.9997878412794807f * rad2deg;
    const float p3 = -0.3258083974640975f * rad2deg;
    const float p5 = 0.1555786518463281f * rad2deg;
    const float p7 = -0.04432655554792128f * rad2deg;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* orb_impl ctor (orb_impl.cc:51-66) evaluates to this table for half patch size 15. */
static const int k_u_max[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

/* orb_impl::ic_angle (orb_impl.cc:68-91) on the UN-blurred level. */
float orc_ic_angle(const uint8_t* img, int stride, int x, int y) {
    int m_01 = 0, m_10 = 0;
    const uint8_t* c = img + (size_t)y * stride + x;
    for (int u = -15; u <= 15; ++u) m_10 += u * c[u];
    for (int v = 1; v <= 15; ++v) {
        int v_sum = 0;
        const int d = k_u_max[v];
        for (int u = -d; u <= d; ++u) {
            const int vp = c[u + v * stride], vm = c[u - v * stride];
            v_sum += vp - vm;
            m_10 += u * (vp + vm);
        }
        m_01 += v * v_sum;
    }
    return orc_fast_atan2((float)m_01, (float)m_10);
}

/* util::cos / util::sin (util/trigonometric.h:11-46). */
static inline float poly_cos(float v) {
    const float c1 = 0.99940307f, c2 = -0.49558072f, c3 = 0.03679168f;
    const float v2 = v * v;
    return c1 + v2 * (c2 + c3 * v2);
}
float orc_util_cos(float v) {
    const float PI = 3.14159265358979f, PI_2 = PI / 2.0f, TWO_PI = 2.0f * PI, INV_TWO_PI = 1.0f / TWO_PI,
                THREE_PI_2 = 3.0f * PI_2;
    v = v - floor_f(v * INV_TWO_PI) * TWO_PI;
    v = (0.0f < v) ? v : -v;
    if (v < PI_2) return poly_cos(v);
    if (v < PI) return -poly_cos(PI - v);
    if (v < THREE_PI_2) return -poly_cos(v - PI);
    return poly_cos(TWO_PI - v);
}
float orc_util_sin(float v) {
    const float PI = 3.14159265358979f, PI_2 = PI / 2.0f;
    return orc_util_cos(PI_2 - v);
}

/* orb_impl::compute_orb_descriptor (orb_impl.cc:93-154): blurred level, fp32 rotate without FMA, cvRound. */
void orc_rbrief(const uint8_t* blurred, int stride, float x, float y, float angle_deg, uint8_t* desc32) {
    const float angle = (float)(angle_deg * 3.14159265358979323846 / 180.0);
    const float ca = orc_util_cos(angle), sa = orc_util_sin(angle);
    const uint8_t* c = blurred + (size_t)round_half_even_f(y) * stride + round_half_even_f(x);
    for (int i = 0; i < 32; ++i) {
        int val = 0;
        for (int b = 0; b < 8; ++b) {
            const int8_t* p = k_pattern + (i * 8 + b) * 4;
            const float x0 = p[0], y0 = p[1], x1 = p[2], y1 = p[3];
            const int r0 = round_half_even_f(x0 * sa + y0 * ca), c0 = round_half_even_f(x0 * ca - y0 * sa);
            const int r1 = round_half_even_f(x1 * sa + y1 * ca), c1 = round_half_even_f(x1 * ca - y1 * sa);
            val |= (c[r0 * stride + c0] < c[r1 * stride + c1]) << b;
        }
        desc32[i] = (uint8_t)val;
    }
}

/* orb_extractor::create_rectangle_mask (orb_extractor.cc:138-151).  cv::rectangle(filled, LINE_AA) zeroes the
 * inclusive rectangle clipped to the image; the anti-aliased fringe is non-zero and the extractor only tests
 * "== 0" (orb_extractor.cc:169), so only the zero set is restated (checked against cv2 in tests). */
void orc_rect_mask(int cols, int rows, const float* rects, int n_rects, uint8_t* mask, int mstride) {
    for (int y = 0; y < rows; ++y) memset(mask + (size_t)y * mstride, 255, cols);
    for (int r = 0; r < n_rects; ++r) {
        const float* q = rects + 4 * r;
        const unsigned x_min = (unsigned)roundf(cols * q[0]), x_max = (unsigned)roundf(cols * q[1]);
        const unsigned y_min = (unsigned)roundf(rows * q[2]), y_max = (unsigned)roundf(rows * q[3]);
        for (unsigned y = y_min; y <= y_max && y < (unsigned)rows; ++y)
            for (unsigned x = x_min; x <= x_max && x < (unsigned)cols; ++x) mask[(size_t)y * mstride + x] = 0;
    }
}

/* ------------------------------------------------------------------------------------------------
 * orb_extractor::extract and helpers (orb_extractor.cc:28-136, 153-345), non-OpenMP order.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    float x, y, response;
} cand_t;

static inline int mask_is_zero(const uint8_t* mask, int mstride, unsigned y, unsigned x, float sf) {
    /* orb_extractor.cc:168-170: mask.at<uchar>(y * scale_factor, x * scale_factor), float product truncated */
    const int r = (int)((float)y * sf), c = (int)((float)x * sf);
    return mask[(size_t)r * mstride + c] == 0;
}

int orc_orb_extract(const uint8_t* img, int w, int h, int stride, const uint8_t* mask, int mask_stride,
                    const orc_orb_config_t* cfg, orc_keypoint_t* kps, uint8_t* descs, int cap, int* level_counts,
                    int* raw_counts, uint8_t** pyramid_out) {
    const int nl = cfg->num_levels;
    const unsigned border = 19, overlap = 6, cell_size = 64;
    const unsigned min_area_sqrt = (unsigned)sqrt((double)cfg->min_area); /* orb_extractor.cc:20 */
    float* sf = (float*)malloc(sizeof(float) * nl);
    orc_scale_factors(cfg->scale_factor, nl, sf, NULL, NULL, NULL);
    uint8_t** pyr = (uint8_t**)calloc(nl, sizeof(uint8_t*));
    int* lw = (int*)malloc(sizeof(int) * nl);
    int* lh = (int*)malloc(sizeof(int) * nl);
    int total = 0, overflow = 0;
    if (w <= 0 || h <= 0) goto done;

    /* compute_image_pyramid (orb_extractor.cc:153-162) */
    lw[0] = w;
    lh[0] = h;
    pyr[0] = (uint8_t*)malloc((size_t)w * h);
    for (int y = 0; y < h; ++y) memcpy(pyr[0] + (size_t)y * w, img + (size_t)y * stride, w);
    for (int l = 1; l < nl; ++l) {
        orc_level_size(w, h, sf[l], &lw[l], &lh[l]);
        pyr[l] = (uint8_t*)malloc((size_t)lw[l] * lh[l]);
        orc_resize_linear_u8(pyr[l - 1], lw[l - 1], lh[l - 1], lw[l - 1], pyr[l], lw[l], lh[l], lw[l]);
    }
    if (pyramid_out)
        for (int l = 0; l < nl; ++l)
            if (pyramid_out[l]) memcpy(pyramid_out[l], pyr[l], (size_t)lw[l] * lh[l]);

    for (int l = 0; l < nl; ++l) {
        const float scale_factor = sf[l];
        const int W = lw[l], H = lh[l];
        int n_level = 0, n_raw = 0;
        if (level_counts) level_counts[l] = 0;
        if (raw_counts) raw_counts[l] = 0;
        if (W <= 2 * (int)border || H <= 2 * (int)border) continue;
        const unsigned max_border_x = W - border, max_border_y = H - border;
        const unsigned width = max_border_x - border, height = max_border_y - border;
        const unsigned num_cols = width / cell_size + 1, num_rows = height / cell_size + 1;

        /* compute_fast_keypoints (orb_extractor.cc:164-265) */
        size_t cand_cap = 4096;
        cand_t* cand = (cand_t*)malloc(sizeof(cand_t) * cand_cap);
        int16_t cx[70 * 70], cy[70 * 70];
        uint8_t cs[70 * 70];
        for (unsigned i = 0; i < num_rows; ++i) {
            const unsigned min_y = border + i * cell_size;
            if (max_border_y - overlap <= min_y) continue;
            unsigned max_y = min_y + cell_size + overlap;
            if (max_border_y < max_y) max_y = max_border_y;
            for (unsigned j = 0; j < num_cols; ++j) {
                const unsigned min_x = border + j * cell_size;
                if (max_border_x - overlap <= min_x) continue;
                unsigned max_x = min_x + cell_size + overlap;
                if (max_border_x < max_x) max_x = max_border_x;
                if (mask) {
                    if (mask_is_zero(mask, mask_stride, min_y, min_x, scale_factor) || mask_is_zero(mask, mask_stride, max_y, min_x, scale_factor)
                        || mask_is_zero(mask, mask_stride, min_y, max_x, scale_factor) || mask_is_zero(mask, mask_stride, max_y, max_x, scale_factor))
                        continue;
                }
                const uint8_t* sub = pyr[l] + (size_t)min_y * W + min_x;
                int n = orc_fast9_16_nms(sub, W, max_x - min_x, max_y - min_y, cfg->ini_fast_thr, cx, cy, cs, 70 * 70);
                if (n == 0) n = orc_fast9_16_nms(sub, W, max_x - min_x, max_y - min_y, cfg->min_fast_thr, cx, cy, cs, 70 * 70);
                for (int k = 0; k < n; ++k) {
                    const float px = (float)cx[k] + (float)(j * cell_size), py = (float)cy[k] + (float)(i * cell_size);
                    if (mask) {
                        const unsigned my = (unsigned)((float)border + py), mx = (unsigned)((float)border + px);
                        if (mask_is_zero(mask, mask_stride, my, mx, scale_factor)) continue;
                    }
                    if ((size_t)n_raw == cand_cap) {
                        cand_cap *= 2;
                        cand = (cand_t*)realloc(cand, sizeof(cand_t) * cand_cap);
                    }
                    cand[n_raw].x = px;
                    cand[n_raw].y = py;
                    cand[n_raw].response = (float)cs[k];
                    ++n_raw;
                }
            }
        }
        if (raw_counts) raw_counts[l] = n_raw;

        /* distribute_keypoints (orb_extractor.cc:289-329): unsigned/float division, then widened */
        const double scaled_min_area_sqrt = (double)((float)min_area_sqrt / scale_factor);
        const int span_x = (int)max_border_x - (int)border, span_y = (int)max_border_y - (int)border;
        const unsigned num_x_grid = (unsigned)ceil(span_x / scaled_min_area_sqrt);
        const unsigned num_y_grid = (unsigned)ceil(span_y / scaled_min_area_sqrt);
        const double delta_x = (double)span_x / num_x_grid, delta_y = (double)span_y / num_y_grid;
        const size_t n_grid = (size_t)num_x_grid * num_y_grid;
        int* winner = (int*)malloc(sizeof(int) * (n_grid ? n_grid : 1));
        for (size_t g = 0; g < n_grid; ++g) winner[g] = -1;
        for (int k = 0; k < n_raw; ++k) {
            const unsigned ix = (unsigned)(cand[k].x / delta_x), iy = (unsigned)(cand[k].y / delta_y);
            const size_t idx = ix + (size_t)iy * num_x_grid;
            if (winner[idx] < 0 || cand[k].response > cand[winner[idx]].response) winner[idx] = k;
        }

        /* orb_extractor.cc:273-285 + compute_orientation; then blur + descriptors + scale correction (:94-129) */
        const unsigned scaled_patch_size = (unsigned)(31 * scale_factor);
        int first = total;
        for (size_t g = 0; g < n_grid; ++g) {
            if (winner[g] < 0) continue;
            if (total >= cap) {
                overflow = 1;
                break;
            }
            orc_keypoint_t* kp = &kps[total];
            kp->x = cand[winner[g]].x + (float)border;
            kp->y = cand[winner[g]].y + (float)border;
            kp->response = cand[winner[g]].response;
            kp->octave = l;
            kp->size = (float)scaled_patch_size;
            kp->angle = orc_ic_angle(pyr[l], W, round_half_even_f(kp->x), round_half_even_f(kp->y));
            ++total;
            ++n_level;
        }
        free(winner);
        free(cand);
        if (overflow) break;
        if (level_counts) level_counts[l] = n_level;
        if (n_level == 0) continue;
        uint8_t* blurred = (uint8_t*)malloc((size_t)W * H);
        orc_gaussian7_s2_u8(pyr[l], W, H, W, blurred, W);
        for (int k = first; k < total; ++k) {
            orc_rbrief(blurred, W, kps[k].x, kps[k].y, kps[k].angle, descs + (size_t)k * 32);
            if (l > 0) { /* correct_keypoint_scale (orb_extractor.cc:337-345) */
                kps[k].x *= scale_factor;
                kps[k].y *= scale_factor;
            }
        }
        free(blurred);
    }
done:
    for (int l = 0; l < nl; ++l) free(pyr[l]);
    free(pyr);
    free(lw);
    free(lh);
    free(sf);
    return overflow ? -1 : total;
}

/* util::convert_to_grayscale (src/stella_vslam/util/image_converter.cc:8-39): cv::cvtColor(COLOR_{RGB,BGR,RGBA,BGRA}2GRAY) on 8-bit
 * images (EXT: OpenCV imgproc, RGB2Gray<uchar>: 15-bit coefficients R 9798, G 19235, B 3735, rounding 1 << 14; pinned against cv2 4.13
 * by tests/test_gray.py).  channels: 3 or 4; rgb_order != 0: the first channel is R. */
void orc_convert_to_grayscale(const uint8_t* src, int w, int h, int stride, int channels, int rgb_order, uint8_t* dst, int dstride) {
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const uint8_t* p = src + (size_t)y * stride + (size_t)x * channels;
            const int c0 = p[0], g = p[1], c2 = p[2];
            const int r = rgb_order ? c0 : c2, b = rgb_order ? c2 : c0;
            dst[(size_t)y * dstride + x] = (uint8_t)((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15);
        }
}
