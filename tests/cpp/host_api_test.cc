// Compiles against include/b200vslam.hpp only (no OpenCV/Eigen).  Without a GPU it checks the "fails loudly" contract;
// with a GPU it runs one extract + match + tiny BA through the C++ mirror classes.
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
    std::vector<uint8_t> empty_desc;
    ex.extract(nullptr, 0, 0, 0, nullptr, 0, kps, empty_desc);  // empty image: silent return
    return kps.empty() ? 0 : 7;
}
