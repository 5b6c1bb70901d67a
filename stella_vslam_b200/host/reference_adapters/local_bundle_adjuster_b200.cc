// optimize::local_bundle_adjuster_b200: host gather (steps 1-4 of local_bundle_adjuster_g2o.cc:41-304), GPU solve
// (steps 5-7 through b200_lba_solve), host write-back under the map mutex (step 8, :379-430).
#include "stella_vslam/optimize/local_bundle_adjuster_b200.h"

#include "stella_vslam/camera/base.h"
#include "stella_vslam/camera/equirectangular.h"
#include "stella_vslam/camera/fisheye.h"
#include "stella_vslam/camera/perspective.h"
#include "stella_vslam/camera/radial_division.h"
#include "stella_vslam/data/keyframe.h"
#include "stella_vslam/data/landmark.h"
#include "stella_vslam/data/map_database.h"
#include "stella_vslam/data/marker.h"
#include "stella_vslam/feature/orb_params.h"

#include <map>
#include <mutex>
#include <unordered_map>

#include <stdexcept>

#include "b200vslam.h"

namespace stella_vslam {
namespace optimize {

local_bundle_adjuster_b200::local_bundle_adjuster_b200(const YAML::Node& yaml_node, unsigned int num_first_iter, unsigned int num_second_iter)
    : num_first_iter_(num_first_iter), num_second_iter_(num_second_iter),
      use_additional_keyframes_for_monocular_(yaml_node["use_additional_keyframes_for_monocular"].as<bool>(false)) {
    if (b200_lba_create(yaml_node["device"].as<int>(0), &handle_) != B200_OK) throw std::runtime_error(b200_last_error());
}

local_bundle_adjuster_b200::~local_bundle_adjuster_b200() { b200_lba_destroy(handle_); }

namespace {
b200_camera_t to_b200(const camera::base* cam) {
    b200_camera_t c{};
    c.fxb = cam->focal_x_baseline_;
    switch (cam->model_type_) {  // reproj_edge_wrapper.h:64-188: every non-equirectangular model uses the perspective edges
        case camera::model_type_t::Perspective: { auto p = static_cast<const camera::perspective*>(cam); c.fx = p->fx_; c.fy = p->fy_; c.cx = p->cx_; c.cy = p->cy_; break; }
        case camera::model_type_t::Fisheye: { auto p = static_cast<const camera::fisheye*>(cam); c.fx = p->fx_; c.fy = p->fy_; c.cx = p->cx_; c.cy = p->cy_; break; }
        case camera::model_type_t::RadialDivision: { auto p = static_cast<const camera::radial_division*>(cam); c.fx = p->fx_; c.fy = p->fy_; c.cx = p->cx_; c.cy = p->cy_; break; }
        case camera::model_type_t::Equirectangular: { c.model = 1; c.cols = cam->cols_; c.rows = cam->rows_; break; }
    }
    return c;
}
}  // namespace

void local_bundle_adjuster_b200::optimize(data::map_database* map_db, const std::shared_ptr<data::keyframe>& curr_keyfrm,
                                          bool* const force_stop_flag) const {
    // ---- 1. window: covisible keyframes, their landmarks, observers outside the window as fixed keyframes -------------
    std::map<unsigned int, std::shared_ptr<data::keyframe>> local_kfs, fixed_kfs;  // ordered: vertex order is id order here
    bool has_scale = false;
    local_kfs[curr_keyfrm->id_] = curr_keyfrm;
    for (const auto& kf : curr_keyfrm->graph_node_->get_covisibilities()) {
        if (!kf || kf->will_be_erased() || kf->graph_node_->is_spanning_root()) continue;
        if (kf->id_ < map_db->get_fixed_keyframe_id_threshold()) continue;
        local_kfs[kf->id_] = kf;
        has_scale |= kf->camera_->setup_type_ != camera::setup_type_t::Monocular;
    }
    std::map<unsigned int, std::shared_ptr<data::landmark>> local_lms;
    for (const auto& id_kf : local_kfs)
        for (const auto& lm : id_kf.second->get_landmarks())
            if (lm && !lm->will_be_erased()) local_lms.emplace(lm->id_, lm);
    for (const auto& id_lm : local_lms)
        for (const auto& obs : id_lm.second->get_observations()) {
            const auto kf = obs.first.lock();
            if (kf && !kf->will_be_erased() && !local_kfs.count(kf->id_)) fixed_kfs.emplace(kf->id_, kf);
        }
    // markers seen in the local keyframes (:81-96)
    std::map<unsigned int, std::shared_ptr<data::marker>> local_mkrs;
    for (const auto& id_kf : local_kfs)
        for (const auto& mkr : id_kf.second->get_markers())
            if (mkr) local_mkrs.emplace(mkr->id_, mkr);
    // local_bundle_adjuster_g2o.cc:135-147: "ensure that there are always at least two fixed keyframes" moves local_keyfrms.begin() of
    // an UNORDERED map -- whichever keyframe libstdc++'s bucket order yields, i.e. an arbitrary one.  Here the containers are ordered
    // by id, so the OLDEST local keyframes are fixed: one valid instance of the reference's behaviour, and the same one on every run.
    if (use_additional_keyframes_for_monocular_ && !has_scale && fixed_kfs.size() < 2 && local_kfs.size() > 2 - fixed_kfs.size())
        while (fixed_kfs.size() < 2) {
            auto it = local_kfs.begin();
            fixed_kfs.insert(*it);
            local_kfs.erase(it);
        }
    // ---- 2-4. flatten --------------------------------------------------------------------------------------------------
    std::vector<std::shared_ptr<data::keyframe>> kfs;
    std::vector<double> pose_cw;
    std::vector<uint8_t> pose_fixed;
    std::unordered_map<unsigned int, int32_t> kf_index;
    std::vector<b200_camera_t> cams;
    std::unordered_map<const camera::base*, uint8_t> cam_index;
    auto add_kf = [&](const std::shared_ptr<data::keyframe>& kf, bool fixed) {
        kf_index[kf->id_] = static_cast<int32_t>(kfs.size());
        kfs.push_back(kf);
        const Mat44_t T = kf->get_pose_cw();
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) pose_cw.push_back(T(r, c));
        pose_fixed.push_back(fixed);
        if (!cam_index.count(kf->camera_)) {
            cam_index[kf->camera_] = static_cast<uint8_t>(cams.size());
            cams.push_back(to_b200(kf->camera_));
        }
    };
    for (const auto& p : local_kfs) add_kf(p.second, false);
    for (const auto& p : fixed_kfs) add_kf(p.second, true);
    std::vector<std::shared_ptr<data::landmark>> lms;
    std::vector<double> points;
    std::vector<int32_t> e_pose, e_point;
    std::vector<uint8_t> e_cam;
    std::vector<float> e_obs, e_isq, e_delta;
    std::vector<uint8_t> point_fixed, e_robust, e_can_be_outlier;
    const float d2 = std::sqrt(5.99146f), d3 = std::sqrt(7.81473f);  // :204-208
    for (const auto& p : local_lms) {
        const auto& lm = p.second;
        const auto observations = lm->get_observations();
        if (observations.empty()) continue;
        const int32_t li = static_cast<int32_t>(lms.size());
        lms.push_back(lm);
        const Vec3_t pw = lm->get_pos_in_world();
        points.insert(points.end(), {pw(0), pw(1), pw(2)});
        point_fixed.push_back(0);
        for (const auto& obs : observations) {
            const auto kf = obs.first.lock();
            if (!kf || kf->will_be_erased() || !kf_index.count(kf->id_)) continue;
            const auto& kp = kf->frm_obs_.undist_keypts_.at(obs.second);
            e_pose.push_back(kf_index.at(kf->id_));
            e_point.push_back(li);
            e_cam.push_back(cam_index.at(kf->camera_));
            e_obs.insert(e_obs.end(), {kp.pt.x, kp.pt.y, kf->frm_obs_.stereo_x_right_.empty() ? -1.0f : kf->frm_obs_.stereo_x_right_.at(obs.second)});
            e_isq.push_back(kf->orb_params_->inv_level_sigma_sq_.at(kp.octave));
            e_delta.push_back(kf->camera_->setup_type_ == camera::setup_type_t::Monocular ? d2 : d3);
            e_robust.push_back(1);
            e_can_be_outlier.push_back(1);
        }
    }
    // marker corners (:251-304): four point vertices per marker that was initialised before (or is kept fixed), one monocular edge per
    // (corner, observing keyframe of the window) with information = identity, NO Huber kernel (use_huber_loss = false) and never
    // outlier-tested (they live in mkr_reproj_edge_wraps, which steps 6-7 do not visit)
    const size_t n_landmark_points = lms.size();
    std::vector<std::pair<std::shared_ptr<data::marker>, size_t>> mkr_points;  // marker, index of its first corner in `points`
    for (const auto& id_mkr : local_mkrs) {
        const auto& mkr = id_mkr.second;
        if (!mkr->keep_fixed_ && !mkr->initialized_before_) continue;
        mkr_points.emplace_back(mkr, points.size() / 3);
        for (unsigned int corner_idx = 0; corner_idx < mkr->corners_pos_w_.size(); ++corner_idx) {
            const int32_t pi = static_cast<int32_t>(points.size() / 3);
            const Vec3_t pw = mkr->corners_pos_w_[corner_idx];
            points.insert(points.end(), {pw(0), pw(1), pw(2)});
            point_fixed.push_back(mkr->keep_fixed_ ? 1 : 0);
            for (const auto& id_kf : mkr->observations_) {
                const auto& kf = id_kf.second;
                if (!kf || kf->will_be_erased() || !kf_index.count(kf->id_)) continue;
                const auto& undist_pt = kf->markers_2d_.at(mkr->id_).undist_corners_.at(corner_idx);
                e_pose.push_back(kf_index.at(kf->id_));
                e_point.push_back(pi);
                e_cam.push_back(cam_index.at(kf->camera_));
                e_obs.insert(e_obs.end(), {undist_pt.x, undist_pt.y, -1.0f});
                e_isq.push_back(1.0f);
                e_delta.push_back(mkr->keep_fixed_ ? 0.0f : d2);
                e_robust.push_back(0);
                e_can_be_outlier.push_back(0);
            }
        }
    }
    b200_lba_problem_t prob{};
    prob.n_poses = static_cast<int32_t>(kfs.size()); prob.n_points = static_cast<int32_t>(points.size() / 3);
    prob.point_fixed = point_fixed.data(); prob.e_robust = e_robust.data(); prob.e_can_be_outlier = e_can_be_outlier.data();
    prob.n_edges = static_cast<int32_t>(e_pose.size()); prob.n_cams = static_cast<int32_t>(cams.size());
    prob.pose_cw = pose_cw.data(); prob.pose_fixed = pose_fixed.data(); prob.points = points.data();
    prob.e_pose = e_pose.data(); prob.e_point = e_point.data(); prob.e_cam = e_cam.data(); prob.e_obs = e_obs.data();
    prob.e_inv_sigma_sq = e_isq.data(); prob.e_delta = e_delta.data(); prob.cams = cams.data();
    // ---- 5-7. GPU -------------------------------------------------------------------------------------------------------
    std::vector<double> pose_out(pose_cw.size()), points_out(points.size());
    std::vector<uint8_t> outlier(e_pose.size());
    static_assert(sizeof(bool) == 1, "the abort flag is passed through as a byte");
    const int rc = b200_lba_solve(handle_, &prob, num_first_iter_, num_second_iter_, reinterpret_cast<volatile uint8_t*>(force_stop_flag),
                                  pose_out.data(), points_out.data(), outlier.data(), nullptr);
    if (rc == B200_ERR_ABORTED) return;  // :308-310
    if (rc != B200_OK) throw std::runtime_error(b200_last_error());
    // ---- 8. write back (:379-430) -----------------------------------------------------------------------------------------
    std::lock_guard<std::mutex> lock(data::map_database::mtx_database_);
    for (size_t e = 0; e < outlier.size(); ++e) {
        if (!outlier[e]) continue;
        if (static_cast<size_t>(e_point[e]) >= n_landmark_points) continue;  // (marker corners are never flagged)
        const auto& kf = kfs[e_pose[e]];
        const auto& lm = lms[e_point[e]];
        if (lm->will_be_erased()) continue;
        kf->erase_landmark(lm);
        lm->erase_observation(map_db, kf);
        if (!lm->will_be_erased()) {
            lm->compute_descriptor();
            lm->update_mean_normal_and_obs_scale_variance();
        }
    }
    for (size_t k = 0; k < kfs.size(); ++k) {
        if (pose_fixed[k]) continue;
        Mat44_t T;
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) T(r, c) = pose_out[16 * k + 4 * r + c];
        kfs[k]->set_pose_cw(T);
    }
    for (size_t l = 0; l < lms.size(); ++l) {
        if (lms[l]->will_be_erased()) continue;
        lms[l]->set_pos_in_world(Vec3_t(points_out[3 * l], points_out[3 * l + 1], points_out[3 * l + 2]));
        lms[l]->update_mean_normal_and_obs_scale_variance();
    }
    for (const auto& mp : mkr_points) {  // :411-428 (a marker has its vertices iff it is listed in mkr_points: the mkr_has_vtx guard)
        const auto& mkr = mp.first;
        if (mkr->keep_fixed_ || !mkr->initialized_before_) continue;
        for (size_t corner_idx = 0; corner_idx < 4 && corner_idx < mkr->corners_pos_w_.size(); ++corner_idx) {
            const size_t pi = mp.second + corner_idx;
            mkr->corners_pos_w_[corner_idx] = Vec3_t(points_out[3 * pi], points_out[3 * pi + 1], points_out[3 * pi + 2]);
        }
    }
}

}  // namespace optimize
}  // namespace stella_vslam
