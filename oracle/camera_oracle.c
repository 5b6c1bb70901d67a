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
