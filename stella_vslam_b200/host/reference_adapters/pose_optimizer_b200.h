// New backend next to optimize/pose_optimizer_g2o.h / _gtsam.h; selected by `Tracking: backend: "b200"` through a branch added to
// optimize/pose_optimizer_factory.h exactly like the "gtsam" one (:27-43):
//
//     else if (backend == "b200") {
//     #ifdef USE_B200
//         YAML::Node node = util::yaml_optional_ref(yaml_node, "b200");
//         return std::unique_ptr<pose_optimizer>(new pose_optimizer_b200(node["num_trials_robust"].as<unsigned int>(2),
//                                                                        node["num_trials"].as<unsigned int>(2),
//                                                                        node["num_each_iter"].as<unsigned int>(10)));
//     #else
//         throw std::runtime_error("b200 backend is not available");
//     #endif
//     }
#ifndef STELLA_VSLAM_OPTIMIZE_POSE_OPTIMIZER_B200_H
#define STELLA_VSLAM_OPTIMIZE_POSE_OPTIMIZER_B200_H

#include <memory>
#include <vector>

// (optimize/pose_optimizer.h names data::landmark without declaring it; the reference's own backends are only ever included after a
//  header that does)
namespace stella_vslam {
namespace data {
class landmark;
}  // namespace data
}  // namespace stella_vslam

#include "stella_vslam/optimize/pose_optimizer.h"

struct b200_lba_s;

namespace stella_vslam {
namespace optimize {

class pose_optimizer_b200 : public pose_optimizer {
public:
    explicit pose_optimizer_b200(unsigned int num_trials_robust = 2, unsigned int num_trials = 2, unsigned int num_each_iter = 10);
    ~pose_optimizer_b200();  // (the reference interface declares no virtual destructor)
    unsigned int optimize(const data::frame& frm, Mat44_t& optimized_pose, std::vector<bool>& outlier_flags) const override;
    unsigned int optimize(const data::keyframe* keyfrm, Mat44_t& optimized_pose, std::vector<bool>& outlier_flags) const override;
    unsigned int optimize(const Mat44_t& cam_pose_cw, const data::frame_observation& frm_obs, const feature::orb_params* orb_params,
                          const camera::base* camera, const std::vector<std::shared_ptr<data::landmark>>& landmarks, Mat44_t& optimized_pose,
                          std::vector<bool>& outlier_flags) const override;

private:
    const unsigned int num_trials_robust_, num_trials_, num_each_iter_;
    b200_lba_s* handle_ = nullptr;
};

}  // namespace optimize
}  // namespace stella_vslam
#endif
