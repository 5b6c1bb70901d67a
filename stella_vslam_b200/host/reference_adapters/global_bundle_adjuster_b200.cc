// optimize::global_bundle_adjuster on the GPU: this translation unit REPLACES src/stella_vslam/optimize/global_bundle_adjuster.cc when
// USE_B200 is set (the class is concrete -- loop_bundle_adjuster.cc:54 and initializer.cc:281 construct it directly -- so the drop-in
// is link-time, like the ORB extractor).  Same gather and write-back as the reference (global_bundle_adjuster.cc:26-192, 201-420); the
// g2o optimiser between them becomes one b200_global_ba_solve call.
#include "stella_vslam/camera/base.h"
#include "stella_vslam/camera/equirectangular.h"
#include "stella_vslam/camera/fisheye.h"
#include "stella_vslam/camera/perspective.h"
#include "stella_vslam/camera/radial_division.h"
#include "stella_vslam/data/keyframe.h"
#include "stella_vslam/data/landmark.h"
#include "stella_vslam/data/marker.h"
#include "stella_vslam/data/map_database.h"
#include "stella_vslam/feature/orb_params.h"
#include "stella_vslam/optimize/global_bundle_adjuster.h"

#include <cmath>
#include <stdexcept>
#include <unordered_map>
#include <unordered_set>

#include "b200vslam.h"

namespace stella_vslam {
namespace optimize {

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

// The flattened problem of optimize_impl (:26-192) and the result of its solve.
struct flat_problem {
    std::vector<std::shared_ptr<data::keyframe>> kfs;  // vertex order
    std::unordered_map<unsigned int, int32_t> kf_index;
    std::vector<double> pose_cw, points;
    std::vector<uint8_t> pose_fixed, point_fixed, e_cam, e_robust;
    std::vector<b200_camera_t> cams;
    std::unordered_map<const camera::base*, uint8_t> cam_index;
    std::vector<int32_t> e_pose, e_point;
    std::vector<float> e_obs, e_isq, e_delta;
    std::vector<int32_t> lm_point;                                    // per entry of `lms`: its point row or -1 (is_optimized_lm, :136-139)
    std::vector<std::pair<std::shared_ptr<data::marker>, size_t>> mkr_points;  // markers that got vertices (mkr_has_vtx) and their first row
    std::vector<double> pose_out, points_out;
};

// returns false if the caller aborted
bool solve_flat(flat_problem& fp, const std::vector<std::shared_ptr<data::keyframe>>& keyfrms, const std::vector<std::shared_ptr<data::landmark>>& lms,
                const std::vector<std::shared_ptr<data::marker>>& markers, unsigned int num_iter, bool use_huber_kernel, bool fix_markers,
                double gain_threshold, bool* const force_stop_flag) {
    // 3. keyframe vertices (:72-84): every keyframe is free except the spanning root
    for (const auto& keyfrm : keyfrms) {
        if (!keyfrm || keyfrm->will_be_erased()) continue;
        fp.kf_index[keyfrm->id_] = static_cast<int32_t>(fp.kfs.size());
        fp.kfs.push_back(keyfrm);
        const Mat44_t T = keyfrm->get_pose_cw();
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) fp.pose_cw.push_back(T(r, c));
        fp.pose_fixed.push_back(keyfrm->graph_node_->is_spanning_root() ? 1 : 0);
        if (!fp.cam_index.count(keyfrm->camera_)) {
            if (fp.cams.size() >= 255) throw std::runtime_error("b200 global BA: more than 255 distinct cameras");
            fp.cam_index[keyfrm->camera_] = static_cast<uint8_t>(fp.cams.size());
            fp.cams.push_back(to_b200(keyfrm->camera_));
        }
    }
    // 4. landmark vertices and reprojection edges (:98-140)
    const float d2 = std::sqrt(5.99146f), d3 = std::sqrt(7.81473f);
    fp.lm_point.assign(lms.size(), -1);
    for (unsigned int i = 0; i < lms.size(); ++i) {
        const auto& lm = lms.at(i);
        if (!lm || lm->will_be_erased()) continue;
        const int32_t li = static_cast<int32_t>(fp.points.size() / 3);
        unsigned int num_edges = 0;
        for (const auto& obs : lm->get_observations()) {
            const auto keyfrm = obs.first.lock();
            const auto idx = obs.second;
            if (!keyfrm || keyfrm->will_be_erased() || !fp.kf_index.count(keyfrm->id_)) continue;
            const auto& undist_keypt = keyfrm->frm_obs_.undist_keypts_.at(idx);
            fp.e_pose.push_back(fp.kf_index.at(keyfrm->id_));
            fp.e_point.push_back(li);
            fp.e_cam.push_back(fp.cam_index.at(keyfrm->camera_));
            fp.e_obs.insert(fp.e_obs.end(), {undist_keypt.pt.x, undist_keypt.pt.y,
                                             keyfrm->frm_obs_.stereo_x_right_.empty() ? -1.0f : keyfrm->frm_obs_.stereo_x_right_.at(idx)});
            fp.e_isq.push_back(keyfrm->orb_params_->inv_level_sigma_sq_.at(undist_keypt.octave));
            fp.e_delta.push_back(keyfrm->camera_->setup_type_ == camera::setup_type_t::Monocular ? d2 : d3);
            fp.e_robust.push_back(use_huber_kernel ? 1 : 0);
            ++num_edges;
        }
        if (num_edges == 0) continue;  // optimizer.removeVertex(lm_vtx); is_optimized_lm.at(i) = false (:136-139)
        fp.lm_point[i] = li;
        const Vec3_t pw = lm->get_pos_in_world();
        fp.points.insert(fp.points.end(), {pw(0), pw(1), pw(2)});
        fp.point_fixed.push_back(0);
    }
    // marker corners (:143-189)
    for (const auto& mkr : markers) {
        if (!mkr) continue;
        if (!fix_markers && !mkr->keep_fixed_ && !mkr->initialized_before_) continue;
        fp.mkr_points.emplace_back(mkr, fp.points.size() / 3);
        const bool fixed = fix_markers || mkr->keep_fixed_;
        for (unsigned int corner_idx = 0; corner_idx < mkr->corners_pos_w_.size(); ++corner_idx) {
            const int32_t pi = static_cast<int32_t>(fp.points.size() / 3);
            const Vec3_t pw = mkr->corners_pos_w_[corner_idx];
            fp.points.insert(fp.points.end(), {pw(0), pw(1), pw(2)});
            fp.point_fixed.push_back(fixed ? 1 : 0);
            for (const auto& id_keyfrm : mkr->observations_) {
                const auto& keyfrm = id_keyfrm.second;
                if (!keyfrm || keyfrm->will_be_erased() || !fp.kf_index.count(keyfrm->id_)) continue;
                const auto& undist_pt = keyfrm->markers_2d_.at(mkr->id_).undist_corners_.at(corner_idx);
                fp.e_pose.push_back(fp.kf_index.at(keyfrm->id_));
                fp.e_point.push_back(pi);
                fp.e_cam.push_back(fp.cam_index.at(keyfrm->camera_));
                fp.e_obs.insert(fp.e_obs.end(), {undist_pt.x, undist_pt.y, -1.0f});
                fp.e_isq.push_back(1.0f);
                fp.e_delta.push_back(fixed ? 0.0f : d2);
                fp.e_robust.push_back(0);  // use_huber_loss = false (:180-182)
            }
        }
    }
    // 5. optimisation (:187-191)
    b200_lba_problem_t prob{};
    prob.n_poses = static_cast<int32_t>(fp.kfs.size());
    prob.n_points = static_cast<int32_t>(fp.points.size() / 3);
    prob.n_edges = static_cast<int32_t>(fp.e_pose.size());
    prob.n_cams = static_cast<int32_t>(fp.cams.size());
    prob.pose_cw = fp.pose_cw.data();
    prob.pose_fixed = fp.pose_fixed.data();
    prob.points = fp.points.data();
    prob.point_fixed = fp.point_fixed.data();
    prob.e_pose = fp.e_pose.data();
    prob.e_point = fp.e_point.data();
    prob.e_cam = fp.e_cam.data();
    prob.e_obs = fp.e_obs.data();
    prob.e_inv_sigma_sq = fp.e_isq.data();
    prob.e_delta = fp.e_delta.data();
    prob.e_robust = fp.e_robust.data();
    prob.cams = fp.cams.data();
    fp.pose_out.resize(fp.pose_cw.size());
    fp.points_out.resize(fp.points.size());
    static thread_local b200_lba_t handle = nullptr;
    if (!handle && b200_lba_create(0, &handle) != B200_OK) throw std::runtime_error(b200_last_error());
    const int rc = b200_global_ba_solve(handle, &prob, static_cast<int>(num_iter), gain_threshold, reinterpret_cast<volatile uint8_t*>(force_stop_flag),
                                        fp.pose_out.data(), fp.points_out.data(), nullptr);
    if (rc == B200_ERR_ABORTED) return false;
    if (rc != B200_OK) throw std::runtime_error(b200_last_error());
    return true;
}

Mat44_t pose_of(const flat_problem& fp, size_t k) {
    Mat44_t T;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) T(r, c) = fp.pose_out[16 * k + 4 * r + c];
    return T;
}
Vec3_t point_of(const flat_problem& fp, size_t p) { return Vec3_t(fp.points_out[3 * p], fp.points_out[3 * p + 1], fp.points_out[3 * p + 2]); }
}  // namespace

global_bundle_adjuster::global_bundle_adjuster(const unsigned int num_iter, const bool use_huber_kernel, const bool verbose)
    : num_iter_(num_iter), use_huber_kernel_(use_huber_kernel), verbose_(verbose) {}

void global_bundle_adjuster::optimize_for_initialization(const std::vector<std::shared_ptr<data::keyframe>>& keyfrms,
                                                         const std::vector<std::shared_ptr<data::landmark>>& lms,
                                                         const std::vector<std::shared_ptr<data::marker>>& markers, float gain_threshold, bool fix_markers,
                                                         bool* const force_stop_flag) const {
    flat_problem fp;
    if (!solve_flat(fp, keyfrms, lms, markers, num_iter_, use_huber_kernel_, fix_markers, gain_threshold, force_stop_flag)) return;
    if (force_stop_flag && *force_stop_flag) return;  // (:224-226: also after a gain-threshold stop, exactly like the reference)
    for (const auto& keyfrm : keyfrms) {  // :230-237
        if (keyfrm->will_be_erased()) continue;
        keyfrm->set_pose_cw(pose_of(fp, fp.kf_index.at(keyfrm->id_)));
    }
    for (unsigned int i = 0; i < lms.size(); ++i) {  // :239-255
        if (fp.lm_point[i] < 0) continue;
        const auto& lm = lms.at(i);
        if (!lm || lm->will_be_erased()) continue;
        lm->set_pos_in_world(point_of(fp, fp.lm_point[i]));
        lm->update_mean_normal_and_obs_scale_variance();
    }
    for (const auto& mp : fp.mkr_points) {  // :257-270
        const auto& mkr = mp.first;
        if (fix_markers || mkr->keep_fixed_ || !mkr->initialized_before_) continue;
        for (size_t corner_idx = 0; corner_idx < 4 && corner_idx < mkr->corners_pos_w_.size(); ++corner_idx)
            mkr->corners_pos_w_[corner_idx] = point_of(fp, mp.second + corner_idx);
    }
}

bool global_bundle_adjuster::optimize(const std::vector<std::shared_ptr<data::keyframe>>& keyfrms, std::unordered_set<unsigned int>& optimized_keyfrm_ids,
                                      std::unordered_set<unsigned int>& optimized_landmark_ids, std::unordered_set<unsigned int>& optimized_marker_ids,
                                      eigen_alloc_unord_map<unsigned int, Vec3_t>& lm_to_pos_w_after_global_BA,
                                      eigen_alloc_unord_map<unsigned int, Mat44_t>& keyfrm_to_pose_cw_after_global_BA,
                                      eigen_alloc_unord_map<unsigned int, std::array<Vec3_t, 4>>& marker_to_pos_w_after_global_BA,
                                      bool* const force_stop_flag) const {
    // the landmarks and markers of the given keyframes, first occurrence order (:273-312)
    std::unordered_set<unsigned int> already_found_landmark_ids, already_found_marker_ids;
    std::vector<std::shared_ptr<data::landmark>> lms;
    std::vector<std::shared_ptr<data::marker>> markers;
    for (const auto& keyfrm : keyfrms) {
        for (const auto& lm : keyfrm->get_landmarks()) {
            if (!lm || lm->will_be_erased() || !already_found_landmark_ids.insert(lm->id_).second) continue;
            lms.push_back(lm);
        }
    }
    for (const auto& keyfrm : keyfrms) {
        for (const auto& mkr : keyfrm->get_markers()) {
            if (!mkr || !already_found_marker_ids.insert(mkr->id_).second) continue;
            markers.push_back(mkr);
        }
    }
    flat_problem fp;
    if (!solve_flat(fp, keyfrms, lms, markers, num_iter_, use_huber_kernel_, false, 1e-3, force_stop_flag)) return false;  // :340-342
    for (const auto& keyfrm : keyfrms) {  // :348-357
        if (keyfrm->will_be_erased()) continue;
        keyfrm_to_pose_cw_after_global_BA[keyfrm->id_] = pose_of(fp, fp.kf_index.at(keyfrm->id_));
        optimized_keyfrm_ids.insert(keyfrm->id_);
    }
    for (unsigned int i = 0; i < lms.size(); ++i) {  // :359-376
        if (fp.lm_point[i] < 0) continue;
        const auto& lm = lms.at(i);
        if (!lm || lm->will_be_erased()) continue;
        lm_to_pos_w_after_global_BA[lm->id_] = point_of(fp, fp.lm_point[i]);
        optimized_landmark_ids.insert(lm->id_);
    }
    for (const auto& mp : fp.mkr_points) {  // :378-410
        const auto& mkr = mp.first;
        if (mkr->keep_fixed_ || !mkr->initialized_before_) continue;
        bool changed = false;
        std::array<Vec3_t, 4> new_pos_corners;
        for (size_t corner_idx = 0; corner_idx < 4; corner_idx++) {
            const Vec3_t new_pos = point_of(fp, mp.second + corner_idx);
            if (mkr->corners_pos_w_[corner_idx] != new_pos) changed = true;
            new_pos_corners[corner_idx] = new_pos;
        }
        if (!changed) continue;
        optimized_marker_ids.insert(mkr->id_);
        marker_to_pos_w_after_global_BA[mkr->id_] = new_pos_corners;
    }
    return true;
}

}  // namespace optimize
}  // namespace stella_vslam
