// New backend next to optimize/local_bundle_adjuster_g2o.h / _gtsam.h; selected by `Mapping: backend: "b200"` through the
// branch added to optimize/local_bundle_adjuster_factory.h.
#ifndef STELLA_VSLAM_OPTIMIZE_LOCAL_BUNDLE_ADJUSTER_B200_H
#define STELLA_VSLAM_OPTIMIZE_LOCAL_BUNDLE_ADJUSTER_B200_H

#include "stella_vslam/optimize/local_bundle_adjuster.h"

#include <yaml-cpp/yaml.h>

struct b200_lba_s;

namespace stella_vslam {
namespace optimize {

class local_bundle_adjuster_b200 : public local_bundle_adjuster {
public:
    explicit local_bundle_adjuster_b200(const YAML::Node& yaml_node, unsigned int num_first_iter = 5, unsigned int num_second_iter = 10);
    ~local_bundle_adjuster_b200();  // (the reference interface declares no virtual destructor)
    void optimize(data::map_database* map_db, const std::shared_ptr<data::keyframe>& curr_keyfrm, bool* const force_stop_flag) const override;

private:
    const unsigned int num_first_iter_, num_second_iter_;
    const bool use_additional_keyframes_for_monocular_;
    b200_lba_s* handle_ = nullptr;
};

}  // namespace optimize
}  // namespace stella_vslam
#endif
