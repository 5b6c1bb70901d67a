#include "../../../../../stella_vslam_b200/host/reference_adapters/local_bundle_adjuster_b200.h"
