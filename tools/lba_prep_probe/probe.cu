// Host-plan timing harness (no GPU needed): runs b200::lba::solve far enough to build the plan; the first CUDA allocation then fails.
#include "../../stella_vslam_b200/csrc/lba_kernels.cu"
extern "C" int probe_plan(const b200_lba_problem_t* P, double* pose, double* pts, uint8_t* outl) {
    b200::lba::Solver S{};
    return b200::lba::solve(S, P, 5, 10, nullptr, pose, pts, outl, nullptr);
}
