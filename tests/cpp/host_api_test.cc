// Compiles against include/b200vslam.hpp only (no OpenCV/Eigen).  Without a GPU it checks the "fails loudly" contract;
// with a GPU it runs one extract + match + tiny BA through the C++ mirror classes.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "b200vslam.hpp"

int main() {
    std::printf("%s\n", b200_version());
    b200::feature::orb_params prm("ORB setting for test");
    if (prm.scale_factors_.size() != 8 || prm.scale_factors_[1] != 1.2f) return 2;
    if (b200_device_count() == 0) {
        try {
            b200::feature::orb_extractor ex(&prm, 800);
            return 3;  // must not construct without a device
        } catch (const std::runtime_error& e) {
            std::printf("no GPU: %s\n", e.what());
            return std::strstr(e.what(), "no CPU fallback") ? 0 : 4;
        }
    }
    const int w = 640, h = 480;
    std::vector<uint8_t> img((size_t)w * h);
    unsigned s = 12345;
    for (auto& p : img) { s = s * 1664525u + 1013904223u; p = (uint8_t)(s >> 24); }
    for (int y = 100; y < 300; ++y) std::memset(&img[(size_t)y * w + 200], 20, 150);
    b200::feature::orb_extractor ex(&prm, 800);
    std::vector<b200_keypoint_t> kps;
    std::vector<uint8_t> desc;
    ex.extract(img.data(), w, h, w, nullptr, 0, kps, desc);
    std::printf("keypoints %zu\n", kps.size());
    if (kps.empty() || desc.size() != kps.size() * 32) return 5;
    b200::match::robust m(0.8f, true);
    std::vector<std::pair<int, int>> matches;
    m.brute_force_match(desc.data(), &kps[0].angle, sizeof(b200_keypoint_t), (int)kps.size(), desc.data(), &kps[0].angle, sizeof(b200_keypoint_t),
                        nullptr, (int)kps.size(), matches);
    std::printf("self matches %zu\n", matches.size());
    for (const auto& pr : matches)
        if (pr.first != pr.second) return 6;  // identical sets match index to index
    // guided matcher through the C++ mirror: every keypoint reprojects onto itself, so landmark q must take keypoint q
    {
        const int n = (int)kps.size();
        std::vector<float> x(n), y(n), margin(n, 4.0f);
        std::vector<uint8_t> oct(n);
        std::vector<int8_t> lo(n, -1), hi(n, -1);
        std::vector<int32_t> out(n, -2);
        for (int i = 0; i < n; ++i) { x[i] = kps[i].x; y[i] = kps[i].y; oct[i] = (uint8_t)kps[i].octave; }
        b200_guided_problem_t P{};
        P.n_train = n; P.t_x = x.data(); P.t_y = y.data(); P.t_octave = oct.data(); P.t_desc = desc.data();
        P.min_x = 0.f; P.max_x = (float)w; P.min_y = 0.f; P.max_y = (float)h; P.grid_cols = 64; P.grid_rows = 48;
        P.n_queries = n; P.q_desc = desc.data(); P.q_x = x.data(); P.q_y = y.data(); P.q_margin = margin.data();
        P.q_min_level = lo.data(); P.q_max_level = hi.data(); P.match_out = out.data();
        b200::match::projection proj(0.9f, false);
        const unsigned int n_proj = proj.match_frame_and_landmarks(P);
        std::printf("projection matches %u\n", n_proj);
        for (int i = 0; i < n; ++i)
            if (out[i] >= 0 && std::memcmp(&desc[32 * (size_t)out[i]], &desc[32 * (size_t)i], 32) != 0) return 8;  // distance 0 wins
        if (n_proj < (unsigned)n / 2) return 9;
    }
    // stereo through the C++ mirror: the same frame as both eyes gives zero disparity -> x_right = x - 0.01 (stereo.cc:78-82)
    {
        b200::match::stereo st(ex, ex, kps, kps, desc, desc, 40.0f, 0.1f);
        std::vector<float> xr, depth;
        st.compute(xr, depth);
        size_t ok = 0;
        for (size_t i = 0; i < kps.size(); ++i) ok += xr[i] >= 0.f && xr[i] <= kps[i].x && depth[i] > 0.f;
        std::printf("stereo matches %zu\n", ok);
        if (ok < kps.size() / 8) return 10;
    }
    // tracking chain through the C++ mirror: every keypoint back-projected to depth 10 m with the identity pose is a landmark that
    // reprojects onto itself, carries its own descriptor and is observable: the chain must attach (almost) all of them and keep the pose
    {
        const int n = (int)kps.size();
        const double fx = 500.0, fy = 500.0, cx = 320.0, cy = 240.0;
        std::vector<double> pos(3 * (size_t)n), nrm(3 * (size_t)n);
        std::vector<float> lo(n), hi(n);
        for (int i = 0; i < n; ++i) {
            const double z = 10.0, X = (kps[i].x - cx) / fx * z, Y = (kps[i].y - cy) / fy * z, d = std::sqrt(X * X + Y * Y + z * z);
            pos[3 * i] = X; pos[3 * i + 1] = Y; pos[3 * i + 2] = z;
            nrm[3 * i] = X / d; nrm[3 * i + 1] = Y / d; nrm[3 * i + 2] = z / d;
            hi[i] = (float)(d * prm.scale_factors_[kps[i].octave]);
            lo[i] = hi[i] / prm.scale_factors_[7];
        }
        b200_track_params_t tp{};
        tp.cam.model = 0; tp.cam.fx = fx; tp.cam.fy = fy; tp.cam.cx = cx; tp.cam.cy = cy; tp.cam.cols = w; tp.cam.rows = h;
        tp.monocular = 1;
        tp.img_bounds[0] = 0.f; tp.img_bounds[1] = (float)w; tp.img_bounds[2] = 0.f; tp.img_bounds[3] = (float)h;
        tp.grid_cols = 64; tp.grid_rows = 48;
        tp.num_levels = 8; tp.log_scale_factor = std::log(1.2f);
        tp.scale_factors = prm.scale_factors_.data(); tp.inv_level_sigma_sq = prm.inv_level_sigma_sq_.data();
        tp.margin = 5.0f; tp.lowe_ratio = 0.8f; tp.hamming_thr = 100; tp.ray_cos_thr = 0.5f;
        tp.num_trials_robust = 2; tp.num_trials = 2; tp.num_each_iter = 10;
        const double pose[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        std::vector<uint8_t> observable(n), outlier(n);
        std::vector<int32_t> slots(n, -2);
        ex.extract(img.data(), w, h, w, nullptr, 0, kps, desc);  // (the chain reads the extractor's LAST batch)
        b200_track_frame_t f{};
        f.frame = 0; f.pose_cw = pose; f.n_landmarks = n; f.lm_pos_w = pos.data(); f.lm_mean_normal = nrm.data();
        f.lm_min_valid_dist = lo.data(); f.lm_max_valid_dist = hi.data(); f.lm_desc = desc.data(); f.kp_cap = n;
        f.lm_observable = observable.data(); f.kp_landmark_out = slots.data(); f.kp_outlier = outlier.data();
        std::vector<b200_track_frame_t> frames(1, f);
        b200::tracking::local_map_tracker tracker(ex, tp);
        tracker.track(frames);
        std::printf("tracking matches %d inliers %u\n", frames[0].n_matches, frames[0].n_valid);
        if (frames[0].n_keypoints != n || frames[0].n_matches < n / 2 || frames[0].n_valid < (unsigned)(n / 2)) return 11;
        for (int r = 0; r < 3; ++r)
            if (std::fabs(frames[0].pose_cw_out[4 * r + 3]) > 1e-2) return 12;  // exact observations: the pose stays at the identity
    }
    std::vector<uint8_t> empty_desc;
    ex.extract(nullptr, 0, 0, 0, nullptr, 0, kps, empty_desc);  // empty image: silent return
    return kps.empty() ? 0 : 7;
}
