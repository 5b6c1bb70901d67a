// Replacement body for match::area::match_in_consistent_area (src/stella_vslam/match/area.cc:8-98), the initialiser's matcher.
// Guard the original with #ifndef USE_B200.  Queries are the level-0 keypoints of frame 1 at their previously matched positions; the
// "a closer landmark takes the keypoint over" state (area.cc:49-51, 75-87) runs in b200_match_guided (mode AREA).
#include <stdexcept>

#include "stella_vslam/camera/base.h"
#include "stella_vslam/data/frame.h"
#include "stella_vslam/match/area.h"

#include "b200vslam.h"

namespace stella_vslam {
namespace match {

unsigned int area::match_in_consistent_area(data::frame& frm_1, data::frame& frm_2, std::vector<cv::Point2f>& prev_matched_pts,
                                            std::vector<int>& matched_indices_2_in_frm_1, int margin) {
    static thread_local b200_matcher_t h = nullptr;
    if (!h && b200_matcher_create(0, &h) != B200_OK) throw std::runtime_error(b200_last_error());
    const auto& kp1 = frm_1.frm_obs_.undist_keypts_;
    const auto& kp2 = frm_2.frm_obs_.undist_keypts_;
    const size_t n1 = kp1.size(), n2 = kp2.size();
    std::vector<float> tx(n2), ty(n2), ta(n2), qx(n1), qy(n1), qa(n1), qm(n1, static_cast<float>(margin));
    std::vector<uint8_t> toct(n2), valid(n1);
    std::vector<int8_t> lvl(n1, 0);
    for (size_t i = 0; i < n2; ++i) {
        tx[i] = kp2[i].pt.x; ty[i] = kp2[i].pt.y; ta[i] = kp2[i].angle; toct[i] = static_cast<uint8_t>(kp2[i].octave);
    }
    for (size_t i = 0; i < n1; ++i) {
        valid[i] = kp1[i].octave <= 0;  // :20-24: level 0 only
        qx[i] = prev_matched_pts.at(i).x; qy[i] = prev_matched_pts.at(i).y; qa[i] = kp1[i].angle;
    }
    matched_indices_2_in_frm_1.assign(n1, -1);
    b200_guided_problem_t P{};
    P.n_train = static_cast<int32_t>(n2);
    P.t_x = tx.data(); P.t_y = ty.data(); P.t_octave = toct.data(); P.t_angle = ta.data();
    P.t_desc = frm_2.frm_obs_.descriptors_.data;
    const auto& b = frm_2.camera_->img_bounds_;
    P.min_x = b.min_x_; P.max_x = b.max_x_; P.min_y = b.min_y_; P.max_y = b.max_y_;
    P.grid_cols = static_cast<int32_t>(frm_2.frm_obs_.num_grid_cols_);
    P.grid_rows = static_cast<int32_t>(frm_2.frm_obs_.num_grid_rows_);
    P.n_queries = static_cast<int32_t>(n1);
    P.q_desc = frm_1.frm_obs_.descriptors_.data;
    P.q_x = qx.data(); P.q_y = qy.data(); P.q_margin = qm.data();
    P.q_min_level = lvl.data(); P.q_max_level = lvl.data();  // get_keypoints_in_cell(..., scale_level_1, scale_level_1) with level 0 (:26-27)
    P.q_angle = qa.data(); P.q_valid = valid.data();
    P.match_out = matched_indices_2_in_frm_1.data();
    if (b200_match_guided(h, 1, &P, B200_GUIDED_AREA, HAMMING_DIST_THR_LOW, lowe_ratio_, check_orientation_ ? 1 : 0, 0) != B200_OK)
        throw std::runtime_error(b200_last_error());
    for (size_t i = 0; i < n1; ++i)  // :89-95
        if (0 <= matched_indices_2_in_frm_1[i]) prev_matched_pts.at(i) = kp2.at(matched_indices_2_in_frm_1[i]).pt;
    return static_cast<unsigned int>(P.n_matches);
}

}  // namespace match
}  // namespace stella_vslam
