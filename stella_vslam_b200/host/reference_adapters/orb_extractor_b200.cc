// Drop-in replacement for src/stella_vslam/feature/orb_extractor.cc (same header, same public members): link this
// translation unit instead of the original one.  system.cc:98-100 constructs the extractor directly, so there is no
// factory to extend; the class keeps its name and signature (feature/orb_extractor.h:46-122).
//
// What moves to the GPU: everything extract() does (orb_extractor.cc:28-136).  What stays: the cv::Mat <-> raw buffer
// glue below and image_pyramid_, which match::stereo reads on the host (system.cc:443) -- it is filled lazily from
// b200_orb_pyramid_level_host so monocular/RGBD runs never pay for the copy.
#include "stella_vslam/feature/orb_extractor.h"

#include <opencv2/core/mat.hpp>

#include <mutex>
#include <unordered_map>

#include <cassert>
#include <stdexcept>

#include "b200vslam.h"

namespace stella_vslam {
namespace feature {

namespace {
// orb_extractor.h has no spare member for the handle; keep it in a side table keyed by `this`.
std::mutex g_mtx;
std::unordered_map<const orb_extractor*, b200_orb_t> g_handles;

b200_orb_t handle_of(const orb_extractor* self, const orb_params* prm, unsigned min_area_sqrt, const std::vector<std::vector<float>>& rects) {
    std::lock_guard<std::mutex> lock(g_mtx);
    auto it = g_handles.find(self);
    if (it != g_handles.end()) return it->second;
    b200_orb_params_t p;
    b200_orb_default_params(&p);
    p.scale_factor = prm->scale_factor_;
    p.num_levels = static_cast<int32_t>(prm->num_levels_);
    p.ini_fast_thr = static_cast<int32_t>(prm->ini_fast_thr_);
    p.min_fast_thr = static_cast<int32_t>(prm->min_fast_thr_);
    p.min_area = min_area_sqrt * min_area_sqrt;  // only (unsigned)sqrt(min_area) is kept by the class (orb_extractor.cc:20)
    std::vector<float> flat;
    for (const auto& r : rects) flat.insert(flat.end(), r.begin(), r.begin() + 4);
    p.n_mask_rects = static_cast<int32_t>(rects.size());
    p.mask_rects = flat.empty() ? nullptr : flat.data();
    b200_orb_t h = nullptr;
    if (b200_orb_create(&p, &h) != B200_OK) throw std::runtime_error(b200_last_error());
    g_handles.emplace(self, h);
    return h;
}
}  // namespace

// the device handle behind an extractor that has extracted at least once (used by stereo_b200.cc); nullptr otherwise
b200_orb_t b200_handle_of(const orb_extractor* self) {
    std::lock_guard<std::mutex> lock(g_mtx);
    auto it = g_handles.find(self);
    return it == g_handles.end() ? nullptr : it->second;
}

orb_extractor::orb_extractor(const orb_params* orb_params, const unsigned int min_area, const descriptor_type desc_type,
                             const std::vector<std::vector<float>>& mask_rects)
    : orb_params_(orb_params), mask_rects_(mask_rects), min_area_sqrt_(std::sqrt(min_area)), desc_type_(desc_type) {
    image_pyramid_.resize(orb_params_->num_levels_);
    if (desc_type_ != descriptor_type::ORB) throw std::runtime_error("the b200 extractor implements descriptor_type::ORB only");
}

void orb_extractor::extract(const cv::_InputArray& in_image, const cv::_InputArray& in_image_mask, std::vector<cv::KeyPoint>& keypts,
                            const cv::_OutputArray& out_descriptors) {
    if (in_image.empty()) return;
    const cv::Mat image = in_image.getMat();
    assert(image.type() == CV_8UC1);
    b200_orb_t h = handle_of(this, orb_params_, min_area_sqrt_, mask_rects_);
    const int cap = b200_orb_max_keypoints(h, image.cols, image.rows);
    std::vector<b200_keypoint_t> kps(cap);
    cv::Mat desc(cap, 32, CV_8U);
    const cv::Mat mask = in_image_mask.empty() ? cv::Mat() : in_image_mask.getMat();
    int32_t n = 0;
    if (b200_orb_extract(h, image.data, image.cols, image.rows, image.step, 0, 1, mask.empty() ? nullptr : mask.data, mask.step, kps.data(),
                         desc.data, cap, &n) != B200_OK)
        throw std::runtime_error(b200_last_error());
    keypts.clear();
    keypts.reserve(n);
    for (int i = 0; i < n; ++i)
        keypts.emplace_back(cv::Point2f(kps[i].x, kps[i].y), kps[i].size, kps[i].angle, kps[i].response, kps[i].octave, -1);
    if (n == 0) out_descriptors.release();
    else desc.rowRange(0, n).copyTo(out_descriptors);
    // image_pyramid_: level 0 aliases the input like the reference (orb_extractor.cc:154); upper levels on demand
    image_pyramid_.at(0) = image;
    for (unsigned int l = 1; l < orb_params_->num_levels_; ++l) {
        int w = 0, hgt = 0;
        b200_orb_level_info(h, static_cast<int>(l), &w, &hgt, nullptr, nullptr);
        image_pyramid_.at(l).create(hgt, w, CV_8UC1);
        b200_orb_pyramid_level_host(h, 0, static_cast<int>(l), image_pyramid_.at(l).data, image_pyramid_.at(l).step);
    }
}

}  // namespace feature
}  // namespace stella_vslam
