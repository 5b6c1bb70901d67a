// Replacement for match::stereo::compute (src/stella_vslam/match/stereo.cc:20-114) when both orb_extractors are the B200 ones
// (orb_extractor_b200.cc): the image pyramids never leave the device.  system.cc:443 constructs match::stereo from
// extractor_left_->image_pyramid_ / extractor_right_->image_pyramid_; with USE_B200 it passes the two extractors instead:
//
//     // system.cc, create_stereo_frame(), replacing :443-447
//     match::stereo_b200(extractor_left_, extractor_right_, keypts_left, keypts_right, frm_obs.descriptors_, descriptors_right,
//                        camera->focal_x_baseline_, camera->true_baseline_).compute(frm_obs.stereo_x_right_, frm_obs.depths_);
//
// orb_extractor_b200.cc exports `b200_orb_t feature::b200_handle_of(const orb_extractor*)` (the handle lives in a side table there).
#include "stella_vslam/feature/orb_extractor.h"
#include "stella_vslam/match/stereo.h"

#include <stdexcept>

#include "b200vslam.h"

namespace stella_vslam {
namespace feature {
b200_orb_t b200_handle_of(const orb_extractor* self);  // orb_extractor_b200.cc
}
namespace match {

class stereo_b200 {
public:
    stereo_b200(const feature::orb_extractor* left, const feature::orb_extractor* right, const std::vector<cv::KeyPoint>& keypts_left,
                const std::vector<cv::KeyPoint>& keypts_right, const cv::Mat& descs_left, const cv::Mat& descs_right, float focal_x_baseline,
                float true_baseline)
        : left_(left), right_(right), keypts_left_(keypts_left), keypts_right_(keypts_right), descs_left_(descs_left), descs_right_(descs_right),
          focal_x_baseline_(focal_x_baseline), true_baseline_(true_baseline) {}

    void compute(std::vector<float>& stereo_x_right, std::vector<float>& depths) const {
        static thread_local b200_matcher_t h = nullptr;
        if (!h && b200_matcher_create(0, &h) != B200_OK) throw std::runtime_error(b200_last_error());
        auto flatten = [](const std::vector<cv::KeyPoint>& in) {
            std::vector<b200_keypoint_t> out(in.size());
            for (size_t i = 0; i < in.size(); ++i) out[i] = {in[i].pt.x, in[i].pt.y, in[i].size, in[i].angle, in[i].response, in[i].octave};
            return out;
        };
        const auto kl = flatten(keypts_left_), kr = flatten(keypts_right_);
        stereo_x_right.assign(kl.size(), -1.0f);
        depths.assign(kl.size(), -1.0f);
        int32_t n = 0;
        if (b200_stereo_compute(h, feature::b200_handle_of(left_), 0, feature::b200_handle_of(right_), 0, kl.data(), descs_left_.data, static_cast<int>(kl.size()),
                                kr.data(), descs_right_.data, static_cast<int>(kr.size()), focal_x_baseline_, true_baseline_, stereo_x_right.data(),
                                depths.data(), &n) != B200_OK)
            throw std::runtime_error(b200_last_error());
    }

private:
    const feature::orb_extractor *left_, *right_;
    const std::vector<cv::KeyPoint>&keypts_left_, &keypts_right_;
    const cv::Mat &descs_left_, &descs_right_;
    const float focal_x_baseline_, true_baseline_;
};

}  // namespace match
}  // namespace stella_vslam
