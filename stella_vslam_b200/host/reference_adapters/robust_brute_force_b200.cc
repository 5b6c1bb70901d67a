// Replacement body for match::robust::brute_force_match (src/stella_vslam/match/robust.cc:232-328).  Compile this TU and
// delete the original definition (or guard it with #ifndef USE_B200).  Everything else in robust.cc is unchanged.
#include "stella_vslam/match/robust.h"
#include "stella_vslam/data/frame_observation.h"
#include "stella_vslam/data/keyframe.h"
#include "stella_vslam/data/landmark.h"

#include <stdexcept>

#include "b200vslam.h"

namespace stella_vslam {
namespace match {

unsigned int robust::brute_force_match(const data::frame_observation& frm_obs, const std::shared_ptr<data::keyframe>& keyfrm,
                                       std::vector<std::pair<int, int>>& matches) const {
    static thread_local b200_matcher_t h = nullptr;
    if (!h && b200_matcher_create(0, &h) != B200_OK) throw std::runtime_error(b200_last_error());
    const auto& kp1 = frm_obs.undist_keypts_;
    const auto& kp2 = keyfrm->frm_obs_.undist_keypts_;
    const auto lms_2 = keyfrm->get_landmarks();
    const int32_t n1 = static_cast<int32_t>(kp1.size()), n2 = static_cast<int32_t>(kp2.size()), off = 0;
    std::vector<uint8_t> valid2(n2);
    for (int i = 0; i < n2; ++i) valid2[i] = lms_2.at(i) && !lms_2.at(i)->will_be_erased();  // robust.cc:255-262
    std::vector<int32_t> pairs(2 * static_cast<size_t>(std::max(n1, 1)));
    int32_t n = 0;
    // descriptors are cv::Mat(N, 32, CV_8U) rows; angles are read in place from the cv::KeyPoint arrays
    if (b200_match_bruteforce(h, 1, frm_obs.descriptors_.data, &kp1.data()->angle, sizeof(cv::KeyPoint), &off, &n1,
                              keyfrm->frm_obs_.descriptors_.data, &kp2.data()->angle, sizeof(cv::KeyPoint), valid2.data(), &off, &n2, lowe_ratio_,
                              check_orientation_ ? 1 : 0, pairs.data(), std::max(n1, 1), &n) != B200_OK)
        throw std::runtime_error(b200_last_error());
    matches.clear();
    matches.reserve(n);
    for (int i = 0; i < n; ++i) matches.emplace_back(pairs[2 * i], pairs[2 * i + 1]);
    return static_cast<unsigned int>(n);
}

}  // namespace match
}  // namespace stella_vslam
