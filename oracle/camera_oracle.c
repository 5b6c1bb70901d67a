/*
 * oracle/camera_oracle.c -- TEST INFRASTRUCTURE (see oracle.h).  CPU restatement of the per-keypoint steps between the extractor and
 * the matchers:
 *   camera::perspective::undistort_keypoints      src/stella_vslam/camera/perspective.cc:245-275
 *        = cv::undistortPoints(pts, K, dist(k1,k2,p1,p2,k3), R = I, P = K, TermCriteria(EPS | MAX_ITER, 20, 1e-6))  (EXT: OpenCV
 *          calib3d, cvUndistortPointsInternal; pinned against cv2 4.13 by tests/test_camera_cpu.py)
 *   camera::perspective::convert_point_to_bearing     perspective.cc:117-122
 *   camera::equirectangular::convert_point_to_bearing equirectangular.cc:42-49  (undistortion is the identity there)
 */
#define _GNU_SOURCE
#include <float.h>
#include <math.h>

#include "oracle.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

void orc_undistort_points(const float* xy_in, int n, double fx, double fy, double cx, double cy, const double* dist5, int max_iter, double eps,
                          float* xy_out) {
    const double k0 = dist5[0], k1 = dist5[1], p1 = dist5[2], p2 = dist5[3], k4 = dist5[4]; /* k[0], k[1], k[2], k[3], k[4]; k[5..11] = 0 */
    const double ifx = 1. / fx, ify = 1. / fy;
    for (int i = 0; i < n; ++i) {
        const double u = xy_in[2 * i], v = xy_in[2 * i + 1];
        double x = (u - cx) * ifx, y = (v - cy) * ify;
        const double x0 = x, y0 = y;
        double error = DBL_MAX;
        for (int j = 0;; ++j) {
            if (j >= max_iter) break;
            if (error < eps) break;
            double r2 = x * x + y * y;
            const double icdist = (1 + ((0 * r2 + 0) * r2 + 0) * r2) / (1 + ((k4 * r2 + k1) * r2 + k0) * r2);
            if (icdist < 0) {
                x = (u - cx) * ifx;
                y = (v - cy) * ify;
                break;
            }
            const double deltaX = 2 * p1 * x * y + p2 * (r2 + 2 * x * x) + 0 * r2 + 0 * r2 * r2;
            const double deltaY = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y + 0 * r2 + 0 * r2 * r2;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
            /* reprojection error of the current estimate (criteria.type & EPS) */
            r2 = x * x + y * y;
            const double r4 = r2 * r2, r6 = r4 * r2;
            const double a1 = 2 * x * y, a2 = r2 + 2 * x * x, a3 = r2 + 2 * y * y;
            const double cdist = 1 + k0 * r2 + k1 * r4 + k4 * r6;
            const double icdist2 = 1. / (1 + 0 * r2 + 0 * r4 + 0 * r6);
            const double xd0 = x * cdist * icdist2 + p1 * a1 + p2 * a2 + 0 * r2 + 0 * r4;
            const double yd0 = y * cdist * icdist2 + p1 * a3 + p2 * a1 + 0 * r2 + 0 * r4;
            const double x_proj = xd0 * fx + cx, y_proj = yd0 * fy + cy;
            error = sqrt(pow(x_proj - u, 2) + pow(y_proj - v, 2));
        }
        /* R = I, P = K:  RR = P * R */
        const double xx = fx * x + 0 * y + cx, yy = 0 * x + fy * y + cy, ww = 1. / (0 * x + 0 * y + 1);
        xy_out[2 * i] = (float)(xx * ww);
        xy_out[2 * i + 1] = (float)(yy * ww);
    }
}

void orc_points_to_bearings(const float* xy, int n, int model, double fx, double fy, double cx, double cy, double cols, double rows,
                            double* bearings) {
    for (int i = 0; i < n; ++i) {
        double* b = bearings + 3 * i;
        if (model == 1) { /* equirectangular.cc:42-49 */
            /* cols_ / rows_ are unsigned int (camera/base.h:105-107): float / unsigned is a FLOAT division */
            const double lon = (xy[2 * i] / (float)(unsigned)cols - 0.5) * (2.0 * M_PI);
            const double lat = -(xy[2 * i + 1] / (float)(unsigned)rows - 0.5) * M_PI;
            b[0] = cos(lat) * sin(lon);
            b[1] = -sin(lat);
            b[2] = cos(lat) * cos(lon);
        } else { /* perspective.cc:117-122 */
            const double xn = (xy[2 * i] - cx) / fx, yn = (xy[2 * i + 1] - cy) / fy;
            const double l2 = sqrt(xn * xn + yn * yn + 1.0);
            b[0] = xn / l2;
            b[1] = yn / l2;
            b[2] = 1.0 / l2;
        }
    }
}

/* data::frame::can_observe (src/stella_vslam/data/frame.cc:59-84) for n landmarks of the local map (tracking_module.cc:559-594):
 * camera::perspective / equirectangular::reproject_to_image (perspective.cc:130-148, equirectangular.cc:59-73),
 * landmark::is_inside_in_orb_scale (data/landmark.h:88-92), the viewing-angle test and landmark::predict_scale_level
 * (data/landmark.cc:336-353; float ratio, logf, ceil).  Rt_cw: rot_cw row-major (9) then trans_cw (3); trans_wc: camera centre. */
void orc_can_observe(int model, double fx, double fy, double cx, double cy, double fxb, double cols, double rows, const float* bounds,
                     const double* Rt_cw, const double* trans_wc, int n, const double* pos_w, const double* mean_normal,
                     const float* min_valid_dist, const float* max_valid_dist, float ray_cos_thr, unsigned num_levels, float log_scale_factor,
                     uint8_t* observable, double* reproj, float* x_right, uint32_t* pred_scale_level) {
    for (int i = 0; i < n; ++i) {
        const double* p = pos_w + 3 * i;
        observable[i] = 0;
        reproj[2 * i] = reproj[2 * i + 1] = 0.0;
        x_right[i] = 0.f;
        pred_scale_level[i] = 0;
        const double pcx = Rt_cw[0] * p[0] + Rt_cw[1] * p[1] + Rt_cw[2] * p[2] + Rt_cw[9];
        const double pcy = Rt_cw[3] * p[0] + Rt_cw[4] * p[1] + Rt_cw[5] * p[2] + Rt_cw[10];
        const double pcz = Rt_cw[6] * p[0] + Rt_cw[7] * p[1] + Rt_cw[8] * p[2] + Rt_cw[11];
        double rx, ry;
        float xr;
        if (model == 1) {
            const double nrm = sqrt(pcx * pcx + pcy * pcy + pcz * pcz);
            const double bx = pcx / nrm, by = pcy / nrm, bz = pcz / nrm;
            const double latitude = -asin(by), longitude = atan2(bx, bz);
            rx = cols * (0.5 + longitude / (2.0 * M_PI));
            ry = rows * (0.5 - latitude / M_PI);
            xr = 0.0f;
        } else {
            if (pcz <= 0.0) continue;
            const double z_inv = 1.0 / pcz;
            rx = fx * pcx * z_inv + cx;
            ry = fy * pcy * z_inv + cy;
            xr = (float)(rx - fxb * z_inv);
            if (!(bounds[0] < rx && rx < bounds[1] && bounds[2] < ry && ry < bounds[3])) continue;
        }
        const double vx = p[0] - trans_wc[0], vy = p[1] - trans_wc[1], vz = p[2] - trans_wc[2];
        const double dist = sqrt(vx * vx + vy * vy + vz * vz);
        const float margin_far = (float)1.3, margin_near = (float)(1.0 / 1.3), distf = (float)dist;
        const float max_dist = margin_far * max_valid_dist[i], min_dist = margin_near * min_valid_dist[i];
        if (!(min_dist <= distf && distf <= max_dist)) continue;
        const double* nml = mean_normal + 3 * i;
        const double ray_cos = (vx * nml[0] + vy * nml[1] + vz * nml[2]) / dist;
        if (ray_cos < ray_cos_thr) continue;
        const float ratio = max_valid_dist[i] / distf;
        const int lvl = (int)ceilf(logf(ratio) / log_scale_factor);
        /* predict_scale_level takes num_scale_levels as float (data/landmark.cc:336) */
        const float nl = (float)num_levels;
        uint32_t out;
        if (lvl < 0) out = 0;
        else if (nl <= (float)(unsigned)lvl) out = (uint32_t)(nl - 1);
        else out = (uint32_t)lvl;
        observable[i] = 1;
        reproj[2 * i] = rx;
        reproj[2 * i + 1] = ry;
        x_right[i] = xr;
        pred_scale_level[i] = out;
    }
}

/* data::landmark::update_mean_normal_and_obs_scale_variance (src/stella_vslam/data/landmark.cc:256-311) for n landmarks.
 * Landmark l is observed from the camera centres cam_centers[3 * offsets[l] .. 3 * offsets[l+1]) (keyfrm->get_trans_wc(), in the
 * order the caller walks observations_ -- the reference's own order is that of a pointer-keyed map); ref_center / ref_scale_factor:
 * the reference keyframe's centre and scale_factors_[octave of its keypoint]; inv_scale_factor_last = inv_scale_factors_[levels-1]. */
void orc_landmark_geometry(int n, const double* pos_w, const int32_t* offsets, const double* cam_centers, const double* ref_center,
                           const float* ref_scale_factor, float inv_scale_factor_last, double* mean_normal, float* max_valid_dist,
                           float* min_valid_dist) {
    for (int l = 0; l < n; ++l) {
        const double* p = pos_w + 3 * l;
        double m[3] = {0, 0, 0};
        for (int o = offsets[l]; o < offsets[l + 1]; ++o) {
            const double v[3] = {p[0] - cam_centers[3 * o], p[1] - cam_centers[3 * o + 1], p[2] - cam_centers[3 * o + 2]};
            const double nrm = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
            for (int k = 0; k < 3; ++k) m[k] = m[k] + (nrm > 0 ? v[k] / nrm : v[k]); /* Eigen normalized(): unchanged when the norm is 0 */
        }
        const double mn = sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
        for (int k = 0; k < 3; ++k) mean_normal[3 * l + k] = mn > 0 ? m[k] / mn : m[k];
        const double r[3] = {p[0] - ref_center[3 * l], p[1] - ref_center[3 * l + 1], p[2] - ref_center[3 * l + 2]};
        const double dist = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
        const float mx = (float)(dist * ref_scale_factor[l]);
        max_valid_dist[l] = mx;
        min_valid_dist[l] = mx * inv_scale_factor_last;
    }
}
