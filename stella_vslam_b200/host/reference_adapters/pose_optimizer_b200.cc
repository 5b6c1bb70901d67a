// optimize::pose_optimizer backend on libb200vslam.so: the gather of pose_optimizer_g2o.cc:57-114 (one edge per keypoint with a
// live landmark) stays here, steps 4-5 (:120-175) run in b200_pose_optimize.
#include "stella_vslam/optimize/pose_optimizer_b200.h"
#include "stella_vslam/camera/base.h"
#include "stella_vslam/camera/perspective.h"
#include "stella_vslam/camera/fisheye.h"
#include "stella_vslam/camera/radial_division.h"
#include "stella_vslam/camera/equirectangular.h"
#include "stella_vslam/data/frame.h"
#include "stella_vslam/data/keyframe.h"
#include "stella_vslam/data/landmark.h"

#include <stdexcept>

#include "b200vslam.h"

namespace stella_vslam {
namespace optimize {

pose_optimizer_b200::pose_optimizer_b200(unsigned int num_trials_robust, unsigned int num_trials, unsigned int num_each_iter)
    : num_trials_robust_(num_trials_robust), num_trials_(num_trials), num_each_iter_(num_each_iter) {
    if (b200_lba_create(0, &handle_) != B200_OK) throw std::runtime_error(b200_last_error());
}
pose_optimizer_b200::~pose_optimizer_b200() { b200_lba_destroy(handle_); }

unsigned int pose_optimizer_b200::optimize(const data::frame& frm, Mat44_t& optimized_pose, std::vector<bool>& outlier_flags) const {
    return optimize(frm.get_pose_cw(), frm.frm_obs_, frm.orb_params_, frm.camera_, frm.get_landmarks(), optimized_pose, outlier_flags);
}
unsigned int pose_optimizer_b200::optimize(const data::keyframe* keyfrm, Mat44_t& optimized_pose, std::vector<bool>& outlier_flags) const {
    return optimize(keyfrm->get_pose_cw(), keyfrm->frm_obs_, keyfrm->orb_params_, keyfrm->camera_, keyfrm->get_landmarks(), optimized_pose,
                    outlier_flags);
}

unsigned int pose_optimizer_b200::optimize(const Mat44_t& cam_pose_cw, const data::frame_observation& frm_obs, const feature::orb_params* orb_params,
                                           const camera::base* camera, const std::vector<std::shared_ptr<data::landmark>>& landmarks,
                                           Mat44_t& optimized_pose, std::vector<bool>& outlier_flags) const {
    const unsigned int num_keypts = frm_obs.undist_keypts_.size();
    outlier_flags.assign(num_keypts, false);
    // camera (pose_opt_edge_wrapper.h:60-186: Perspective / Fisheye / RadialDivision share the perspective edges on undistorted keypoints)
    b200_camera_t cam{};
    switch (camera->model_type_) {
        case camera::model_type_t::Perspective: { const auto c = static_cast<const camera::perspective*>(camera); cam = {0, c->fx_, c->fy_, c->cx_, c->cy_, camera->focal_x_baseline_, 0, 0}; break; }
        case camera::model_type_t::Fisheye: { const auto c = static_cast<const camera::fisheye*>(camera); cam = {0, c->fx_, c->fy_, c->cx_, c->cy_, camera->focal_x_baseline_, 0, 0}; break; }
        case camera::model_type_t::RadialDivision: { const auto c = static_cast<const camera::radial_division*>(camera); cam = {0, c->fx_, c->fy_, c->cx_, c->cy_, camera->focal_x_baseline_, 0, 0}; break; }
        case camera::model_type_t::Equirectangular: { const auto c = static_cast<const camera::equirectangular*>(camera); cam = {1, 0, 0, 0, 0, 0, double(c->cols_), double(c->rows_)}; break; }
    }
    const float sqrt_chi_sq = (camera->setup_type_ == camera::setup_type_t::Monocular) ? std::sqrt(5.99146f) : std::sqrt(7.81473f);  // :84-101
    std::vector<double> points;
    std::vector<float> obs, inv_sigma_sq, delta;
    std::vector<int32_t> e_pose, e_point;
    std::vector<uint8_t> e_cam, fixed;
    std::vector<unsigned int> idx_of_edge;
    for (unsigned int idx = 0; idx < num_keypts; ++idx) {
        const auto& lm = landmarks.at(idx);
        if (!lm || lm->will_be_erased()) continue;  // :87-92
        const Vec3_t pos_w = lm->get_pos_in_world();
        const auto& kp = frm_obs.undist_keypts_.at(idx);
        e_point.push_back(static_cast<int32_t>(idx_of_edge.size()));
        idx_of_edge.push_back(idx);
        points.insert(points.end(), {pos_w(0), pos_w(1), pos_w(2)});
        obs.insert(obs.end(), {kp.pt.x, kp.pt.y, frm_obs.stereo_x_right_.empty() ? -1.0f : frm_obs.stereo_x_right_.at(idx)});
        inv_sigma_sq.push_back(orb_params->inv_level_sigma_sq_.at(kp.octave));
        delta.push_back(sqrt_chi_sq);
    }
    const int n = static_cast<int>(idx_of_edge.size());
    e_pose.assign(n, 0);
    e_cam.assign(n, 0);
    fixed.assign(n, 1);
    double pose_in[16], pose_out[16];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) pose_in[4 * r + c] = cam_pose_cw(r, c);
    const uint8_t pose_free = 0;
    b200_lba_problem_t P{};
    P.n_poses = 1; P.n_points = n; P.n_edges = n; P.n_cams = 1;
    P.pose_cw = pose_in; P.pose_fixed = &pose_free; P.points = points.data(); P.point_fixed = fixed.data();
    P.e_pose = e_pose.data(); P.e_point = e_point.data(); P.e_cam = e_cam.data(); P.e_obs = obs.data();
    P.e_inv_sigma_sq = inv_sigma_sq.data(); P.e_delta = delta.data(); P.cams = &cam;
    std::vector<uint8_t> flags(std::max(n, 1));
    uint32_t n_valid = 0;
    if (b200_pose_optimize(handle_, 1, &P, num_trials_robust_, num_trials_, num_each_iter_, pose_out, flags.data(), &n_valid) != B200_OK)
        throw std::runtime_error(b200_last_error());
    if (n < 5) return 0;  // :116-118 (optimized_pose is left untouched)
    for (int e = 0; e < n; ++e) outlier_flags.at(idx_of_edge[e]) = flags[e] != 0;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) optimized_pose(r, c) = pose_out[4 * r + c];
    return n_valid;
}

}  // namespace optimize
}  // namespace stella_vslam
