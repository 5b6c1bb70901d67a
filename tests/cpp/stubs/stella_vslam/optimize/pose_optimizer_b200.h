#include "../../../../../stella_vslam_b200/host/reference_adapters/pose_optimizer_b200.h"
