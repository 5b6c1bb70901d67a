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
    std::vector<uint8_t> empty_desc;
    ex.extract(nullptr, 0, 0, 0, nullptr, 0, kps, empty_desc);  // empty image: silent return
    return kps.empty() ? 0 : 7;
}
