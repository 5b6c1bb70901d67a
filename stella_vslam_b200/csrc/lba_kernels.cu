// lba_kernels.cu -- optimize::local_bundle_adjuster on sm_100a (fp64).
//
// Reference path (paths relative to the reference checkout):
//   local_bundle_adjuster_g2o::optimize steps 5-7     src/stella_vslam/optimize/local_bundle_adjuster_g2o.cc:306-375
//   mono/stereo perspective reprojection edges         optimize/internal/se3/perspective_reproj_edge.h:67-120, 175-239
//   equirectangular reprojection edge                  optimize/internal/se3/equirectangular_reproj_edge.h:64-134
//   edge wrapper (information, Huber delta, mono test) optimize/internal/se3/reproj_edge_wrapper.h:57-268
//   shot_vertex / landmark_vertex oplus                optimize/internal/se3/shot_vertex.h:55-58, internal/landmark_vertex.h:50-53
//   terminate_action (gain threshold 1e-3)             optimize/terminate_action.cc:36-76
// and g2o's published algorithm (tag 20230223_git, not vendored): BaseBinaryEdge::constructQuadraticForm with
// RobustKernelHuber, BlockSolver_6_3 (Schur complement over the landmarks), OptimizationAlgorithmLevenberg.
//
// Layout: edges are sorted by landmark (CSR), so every landmark-side quantity (Hll, bl, Dinv, back-substitution) is a
// contiguous, atomics-free reduction; pose-side blocks are built by one CTA per free keyframe; the Schur complement is a
// block-sparse  Hschur(i,j) = Hpp(i,j) - sum_l Hpl(i,l) Dinv(l) Hpl(j,l)^T  evaluated by one warp per (i,j) block over a
// pair list built once per solve.  Everything is deterministic (fixed reduction orders, no floating-point atomics).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <new>
#include <sched.h>

#include <chrono>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace b200 {
namespace lba {

constexpr double kPi = 3.14159265358979323846;

struct Cam {
    int model;
    double fx, fy, cx, cy, fxb, cols, rows;
};

// per-edge static data (sorted by landmark)
struct EdgeS {
    int pose;        // keyframe index
    int pcol;        // free-pose column or -1
    int point;       // landmark index
    int lcol;        // free-landmark column or -1
    float ox, oy, oxr;
    float inv_sigma_sq;
    float delta;
    unsigned char cam, robust, can_outlier, pad;
};

struct View {
    int K, L, E, Kf, Lf;
    const EdgeS* edges;
    const Cam* cams;
    const int* pt_start;     // L+1 (all landmarks)
    const int* pose_start;   // Kf+1
    const int* pose_edges;   // edge ids grouped by free pose
    unsigned char* level;    // E: 0 active, 1 outlier
    unsigned char* robust;   // E: Huber on/off for the current round
};

// ---------------------------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------------------------
__host__ __device__ inline void quat_to_rot(const double* q, double* R) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
__host__ __device__ inline void rot_to_quat(const double* R, double* q) {
    double t = R[0] + R[4] + R[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (R[7] - R[5]) * t; q[1] = (R[2] - R[6]) * t; q[2] = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[i * 3 + i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(R[i * 3 + i] - R[j * 3 + j] - R[k * 3 + k] + 1.0);
        double qq[4];
        qq[i] = 0.5 * t;
        t = 0.5 / t;
        qq[3] = (R[k * 3 + j] - R[j * 3 + k]) * t;
        qq[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
        qq[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
        q[0] = qq[0]; q[1] = qq[1]; q[2] = qq[2]; q[3] = qq[3];
    }
}
__host__ __device__ inline void quat_normalize(double* q) {  // SE3Quat::normalizeRotation
    if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
    const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}

// Residual of one edge at pose Rt = [R(9) t(3)], landmark p.  dim = 2 (mono / equirect) or 3 (stereo).
__device__ __forceinline__ void edge_residual(const EdgeS& e, const Cam& c, const double* Rt, const double* p, double* err, double* pc) {
    pc[0] = Rt[0] * p[0] + Rt[1] * p[1] + Rt[2] * p[2] + Rt[9];
    pc[1] = Rt[3] * p[0] + Rt[4] * p[1] + Rt[5] * p[2] + Rt[10];
    pc[2] = Rt[6] * p[0] + Rt[7] * p[1] + Rt[8] * p[2] + Rt[11];
    if (c.model == 1) {  // equirectangular_reproj_edge.h:130-134
        const double theta = atan2(pc[0], pc[2]);
        const double phi = -asin(pc[1] / sqrt(pc[0] * pc[0] + pc[1] * pc[1] + pc[2] * pc[2]));
        err[0] = (double)e.ox - c.cols * (0.5 + theta / (2 * kPi));
        err[1] = (double)e.oy - c.rows * (0.5 - phi / kPi);
        err[2] = 0.0;
    } else {  // perspective_reproj_edge.h:118-120, 236-239
        const double rx = c.fx * pc[0] / pc[2] + c.cx;
        err[0] = (double)e.ox - rx;
        err[1] = (double)e.oy - (c.fy * pc[1] / pc[2] + c.cy);
        err[2] = (e.oxr < 0.f) ? 0.0 : (double)e.oxr - (rx - c.fxb / pc[2]);
    }
}

// linearizeOplus: Ji (3 rows x 3, landmark) and Jj (3 rows x 6, pose, rotation first); unused rows are zero.
__device__ __forceinline__ void edge_jacobians(const EdgeS& e, const Cam& c, const double* Rt, const double* pc, double* Ji, double* Jj) {
    const double x = pc[0], y = pc[1], z = pc[2];
#pragma unroll
    for (int i = 0; i < 9; ++i) Ji[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 18; ++i) Jj[i] = 0.0;
    if (c.model == 1) {  // equirectangular_reproj_edge.h:71-128
        const double L = sqrt(x * x + y * y + z * z);
        const double dx[9] = {0, z, -y, 1, 0, 0, Rt[0], Rt[1], Rt[2]};
        const double dy[9] = {-z, 0, x, 0, 1, 0, Rt[3], Rt[4], Rt[5]};
        const double dz[9] = {y, -x, 0, 0, 0, 1, Rt[6], Rt[7], Rt[8]};
        const double k0 = -(c.cols / (2 * kPi)) * (1.0 / (x * x + z * z));
        const double k1 = -(c.rows / kPi) * (1.0 / (L * sqrt(x * x + z * z)));
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const double dL = (1.0 / L) * (x * dx[j] + y * dy[j] + z * dz[j]);
            const double j0 = k0 * (z * dx[j] - x * dz[j]);
            const double j1 = k1 * (L * dy[j] - y * dL);
            if (j < 6) {
                Jj[j] = j0;
                Jj[6 + j] = j1;
            } else {
                Ji[j - 6] = j0;
                Ji[3 + j - 6] = j1;
            }
        }
        return;
    }
    const double fx = c.fx, fy = c.fy, z_sq = z * z;
#pragma unroll
    for (int j = 0; j < 3; ++j) {  // perspective_reproj_edge.h:89-95
        Ji[j] = -fx * Rt[j] / z + fx * x * Rt[6 + j] / z_sq;
        Ji[3 + j] = -fy * Rt[3 + j] / z + fy * y * Rt[6 + j] / z_sq;
    }
    Jj[0] = x * y / z_sq * fx; Jj[1] = -(1.0 + (x * x / z_sq)) * fx; Jj[2] = y / z * fx;
    Jj[3] = -1.0 / z * fx;     Jj[4] = 0.0;                            Jj[5] = x / z_sq * fx;
    Jj[6] = (1.0 + y * y / z_sq) * fy; Jj[7] = -x * y / z_sq * fy; Jj[8] = -x / z * fy;
    Jj[9] = 0.0;                       Jj[10] = -1.0 / z * fy;     Jj[11] = y / z_sq * fy;
    if (e.oxr >= 0.f) {  // perspective_reproj_edge.h:203-205, 221-226
        const double fxb = c.fxb;
#pragma unroll
        for (int j = 0; j < 3; ++j) Ji[6 + j] = Ji[j] - fxb * Rt[6 + j] / z_sq;
        Jj[12] = Jj[0] - fxb * y / z_sq; Jj[13] = Jj[1] + fxb * x / z_sq; Jj[14] = Jj[2];
        Jj[15] = Jj[3];                  Jj[16] = 0.0;                    Jj[17] = Jj[5] - fxb / z_sq;
    }
}

// RobustKernelHuber: rho[1] weight and rho[0] cost
__device__ __forceinline__ double huber_weight(double e2, double delta) { return (e2 <= delta * delta) ? 1.0 : delta / sqrt(e2); }
__device__ __forceinline__ double huber_cost(double e2, double delta) { return (e2 <= delta * delta) ? e2 : 2 * sqrt(e2) * delta - delta * delta; }

__device__ __forceinline__ double block_sum(double v, double* sh) {  // deterministic tree reduction, blockDim.x power of two <= 1024
    const int tid = threadIdx.x;
    sh[tid] = v;
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (tid < s) sh[tid] += sh[tid + s];
        __syncthreads();
    }
    const double r = sh[0];
    __syncthreads();
    return r;
}

// ---------------------------------------------------------------------------------------------------------------
// Levenberg-Marquardt control block.  The whole two-round optimisation runs as ONE CUDA graph whose loops are WHILE conditional
// nodes: the numeric kernels read lambda and the index of the current state from this block, three single-CTA control kernels
// restate g2o's OptimizationAlgorithmLevenberg::solve / SparseOptimizer::optimize / terminate_action bookkeeping and set the
// loop conditions.  By default the host steps the same kernels (one round trip per loop decision); B200_LBA_GRAPH=1 selects the
// graph driver (see b200_lba_create for why it is opt-in).
// ---------------------------------------------------------------------------------------------------------------
struct Dual {
    double* p[2];  // [current / trial] double buffers; LmCtl::cur says which one is current
};
struct LmCtl {
    double lambda, ni, current_chi, last_chi, rho;
    double lambda_init, chi2[2], lambda_final[2];
    int cur, it, iterations, qmax, ok, stop_flag, round, skip_round2;
    int inner_go, outer_go;
    int iters_done[2];
    int launches;
    const volatile int* abort_word;  // mapped host word mirroring the caller's force_stop flag
};
__device__ __forceinline__ bool lm_aborted(const LmCtl* c) { return c->stop_flag || (c->abort_word && *c->abort_word); }

// ---------------------------------------------------------------------------------------------------------------
// LM control kernels (one CTA each).  Sums of the per-CTA partials are taken by thread 0 in index order (staged through
// shared memory), exactly like a host loop over the read-back array would.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kCtlThreads = 256;
__device__ double ordered_sum(const double* __restrict__ p, int n, double* stage) {
    double s = 0.0;
    for (int base = 0; base < n; base += 1024) {
        const int m = min(1024, n - base);
        __syncthreads();
        for (int i = threadIdx.x; i < m; i += blockDim.x) stage[i] = __ldcg(p + base + i);  // L2: partials may come from other CTAs of this launch
        __syncthreads();
        if (threadIdx.x == 0)
            for (int i = 0; i < m; ++i) s += stage[i];
    }
    return s;  // valid in thread 0
}
__device__ __forceinline__ void set_cond(cudaGraphConditionalHandle h, int use_graph, unsigned v) {
    if (use_graph) cudaGraphSetConditional(h, v);
}

// start of SparseOptimizer::optimize(iterations): terminate_action at iteration -1 resets the stop flag (terminate_action.cc:46-51)
__global__ void lm_round_begin_kernel(LmCtl* __restrict__ c, int iterations, int round, cudaGraphConditionalHandle h_outer, int use_graph) {
    if (threadIdx.x) return;
    c->round = round;
    c->iterations = iterations;
    c->iters_done[round] = 0;
    c->launches += 1;
    if (round == 1 && lm_aborted(c)) {  // local_bundle_adjuster_g2o.cc:317-321
        c->skip_round2 = 1;
        c->outer_go = 0;
        set_cond(h_outer, use_graph, 0u);
        return;
    }
    c->stop_flag = 0;
    c->it = 0;
    c->ok = 1;
    c->outer_go = (iterations > 0 && !lm_aborted(c)) ? 1 : 0;
    set_cond(h_outer, use_graph, (unsigned)c->outer_go);
}

// after computeActiveErrors + buildSystem: at the first iteration of a round take the robust chi2 and computeLambdaInit
// (tau * max |H_jj| over all free vertices, tau = 1e-5); arm the trial loop
__device__ void lm_after_build(LmCtl* __restrict__ c, const double* __restrict__ r_chi, int eb, const double* __restrict__ r_diag, int lb,
                               const double* __restrict__ Hpp, int Kf, int n_build_launches, cudaGraphConditionalHandle h_inner, int use_graph,
                               double* stage) {
    const bool first = c->it == 0;
    double chi = 0.0;
    if (first) chi = ordered_sum(r_chi, eb, stage);
    if (threadIdx.x) return;
    if (first) {
        c->current_chi = chi;
        double mx = 0.0;
        for (int i = 0; i < lb; ++i) mx = fmax(mx, __ldcg(r_diag + i));
        for (int p = 0; p < Kf; ++p)
            for (int a = 0; a < 6; ++a) mx = fmax(mx, fabs(__ldcg(Hpp + 36 * (size_t)p + a * 7)));
        c->lambda = 1e-5 * mx;
        c->ni = 2.0;
        if (c->round == 0) c->lambda_init = c->lambda;
    }
    c->qmax = 0;
    c->rho = 0.0;
    c->inner_go = 1;
    c->launches += n_build_launches;
    set_cond(h_inner, use_graph, 1u);
}
__global__ void __launch_bounds__(kCtlThreads) lm_after_build_kernel(LmCtl* __restrict__ c, const double* __restrict__ r_chi, int eb,
                                                                     const double* __restrict__ r_diag, int lb, const double* __restrict__ Hpp,
                                                                     int Kf, int n_build_launches, cudaGraphConditionalHandle h_inner,
                                                                     int use_graph) {
    __shared__ double stage[1024];
    lm_after_build(c, r_chi, eb, r_diag, lb, Hpp, Kf, n_build_launches, h_inner, use_graph, stage);
}

// after one trial (solve, back-substitution, chi2 at the trial state): the accept / reject rule of
// OptimizationAlgorithmLevenberg::solve, and when the trial loop ends the end-of-iteration bookkeeping of
// SparseOptimizer::optimize + terminate_action (terminate_action.cc:52-73)
__device__ void lm_after_trial(LmCtl* __restrict__ c, const double* __restrict__ r_chi, int eb, const double* __restrict__ r_scale, int lb2,
                               const double* __restrict__ r_result, int n_trial_launches, cudaGraphConditionalHandle h_inner,
                               cudaGraphConditionalHandle h_outer, int use_graph, double* stage) {
    const bool ok2 = __ldcg(r_result) != 0.0;
    const double chi_sum = ordered_sum(r_chi, eb, stage);
    const double scale_sum = ordered_sum(r_scale, lb2, stage);
    if (threadIdx.x) return;
    c->launches += n_trial_launches;
    const double temp_chi = ok2 ? chi_sum : 1.7976931348623157e308;
    double rho = c->current_chi - temp_chi;
    double scale = ok2 ? __ldcg(r_result + 1) + scale_sum : 0.0;  // computeScale
    scale += 1e-3;
    rho /= scale;
    bool broke = false;
    if (rho > 0 && isfinite(temp_chi) && ok2) {
        double alpha = 1. - pow(2 * rho - 1, 3.0);
        alpha = fmin(alpha, 2. / 3.);
        c->lambda *= fmax(1. / 3., alpha);
        c->ni = 2.0;
        c->current_chi = temp_chi;
        c->cur ^= 1;  // discardTop: keep the trial state
    } else {
        c->lambda *= c->ni;
        c->ni *= 2.0;  // pop: the current state is untouched
        if (!isfinite(c->lambda)) broke = true;
    }
    if (!broke) c->qmax++;
    c->rho = rho;
    const bool again = !broke && rho < 0 && c->qmax < 10 && !lm_aborted(c);
    c->inner_go = again ? 1 : 0;
    set_cond(h_inner, use_graph, again ? 1u : 0u);
    if (again) return;
    if (c->qmax == 10 || rho == 0 || !isfinite(c->lambda)) c->ok = 0;  // SolverResult::Terminate
    const double chi_now = c->current_chi;
    if (c->it == 0) {
        c->last_chi = chi_now;
    } else {
        const double gain = (c->last_chi - chi_now) / chi_now;
        c->last_chi = chi_now;
        if (gain >= 0 && gain < 1e-3) c->stop_flag = 1;
    }
    c->chi2[c->round] = chi_now;
    c->lambda_final[c->round] = c->lambda;
    c->it++;
    c->iters_done[c->round] = c->it;
    const bool more = c->it < c->iterations && !lm_aborted(c) && c->ok;
    c->outer_go = more ? 1 : 0;
    set_cond(h_outer, use_graph, more ? 1u : 0u);
}
__global__ void __launch_bounds__(kCtlThreads) lm_after_trial_kernel(LmCtl* __restrict__ c, const double* __restrict__ r_chi, int eb,
                                                                     const double* __restrict__ r_scale, int lb2,
                                                                     const double* __restrict__ r_result, int n_trial_launches,
                                                                     cudaGraphConditionalHandle h_inner, cudaGraphConditionalHandle h_outer,
                                                                     int use_graph) {
    __shared__ double stage[1024];
    lm_after_trial(c, r_chi, eb, r_scale, lb2, r_result, n_trial_launches, h_inner, h_outer, use_graph, stage);
}

// "last CTA" election: after every CTA of the launch has published its results, exactly one of them (the last to arrive) sees
// true and may consume them; it re-arms the ticket for the next launch.
__device__ __forceinline__ bool last_cta_arrives(int* ticket, int* smem_flag) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int t = atomicAdd(ticket, 1);
        *smem_flag = (t == (int)gridDim.x - 1);
        if (*smem_flag) *ticket = 0;
    }
    __syncthreads();
    const bool last = *smem_flag != 0;
    if (last) __threadfence();
    return last;
}
// what the tail of a fused launch needs to run the LM bookkeeping
struct CtlTail {
    LmCtl* ctl;
    int* ticket;
    const double *r_chi, *r_diag, *r_scale, *r_result, *Hpp;
    int eb, lb, lb2, Kf, n_launches, use_graph;
    cudaGraphConditionalHandle h_inner, h_outer;
};

// end of a round: a round that ran no iteration still reports the chi2 of its (re-evaluated) state
__global__ void __launch_bounds__(kCtlThreads) lm_round_end_kernel(LmCtl* __restrict__ c, const double* __restrict__ r_chi, int eb, int round) {
    __shared__ double stage[1024];
    if (round == 1 && c->skip_round2) return;
    const double chi = ordered_sum(r_chi, eb, stage);
    if (threadIdx.x) return;
    c->launches += 2;
    if (c->iters_done[round] == 0) c->chi2[round] = chi;
}

// ---------------------------------------------------------------------------------------------------------------
// K1: per edge -- residual, chi2, Huber weight, Hpl block and the landmark-side contribution
//     out: chi[e] (plain chi2, only for active edges), Hpl[18][E], pl[9][E] (6 unique Hll + 3 bl), chi partials
// ---------------------------------------------------------------------------------------------------------------
constexpr int kEdgeThreads = 128;

__global__ void __launch_bounds__(kEdgeThreads) edges_kernel(View v, Dual Rt2, Dual pts2, Dual chi2, double* __restrict__ Hpl,
                                                             double* __restrict__ pl, double* __restrict__ chi_partials, int linearize,
                                                             int on_trial, int* __restrict__ fail_reset, const LmCtl* __restrict__ ctl,
                                                             CtlTail tail) {
    __shared__ double sh[kEdgeThreads];
    const int sidx = (ctl->cur ^ on_trial) & 1;  // on_trial: evaluate the trial state, carrying inactive chi2 over from the current one
    const double* __restrict__ Rt = Rt2.p[sidx];
    const double* __restrict__ pts = pts2.p[sidx];
    double* __restrict__ chi = chi2.p[sidx];
    const double* __restrict__ chi_carry = on_trial ? chi2.p[sidx ^ 1] : nullptr;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (fail_reset && e == 0) *fail_reset = 0;  // last kernel of a trial: re-arm the solver's failure flag
    double cost = 0.0;
    if (e < v.E) {
        const EdgeS ed = v.edges[e];
        const bool active = v.level[e] == 0;
        if (active) {
            const Cam c = v.cams[ed.cam];
            double err[3], pc[3];
            const double* P = pts + 3 * (size_t)ed.point;
            const double* T = Rt + 12 * (size_t)ed.pose;
            edge_residual(ed, c, T, P, err, pc);
            const double w = (double)ed.inv_sigma_sq;
            const double e2 = w * (err[0] * err[0] + err[1] * err[1] + err[2] * err[2]);
            chi[e] = e2;
            const bool rob = v.robust[e] != 0;
            cost = rob ? huber_cost(e2, (double)ed.delta) : e2;
            if (linearize) {
                double Ji[9], Jj[18];
                edge_jacobians(ed, c, T, pc, Ji, Jj);
                const double ww = w * (rob ? huber_weight(e2, (double)ed.delta) : 1.0);
                const bool lfree = ed.lcol >= 0, pfree = ed.pcol >= 0;
                // Hll (upper 6) and bl
                double h[9];
                h[0] = ww * (Ji[0] * Ji[0] + Ji[3] * Ji[3] + Ji[6] * Ji[6]);
                h[1] = ww * (Ji[0] * Ji[1] + Ji[3] * Ji[4] + Ji[6] * Ji[7]);
                h[2] = ww * (Ji[0] * Ji[2] + Ji[3] * Ji[5] + Ji[6] * Ji[8]);
                h[3] = ww * (Ji[1] * Ji[1] + Ji[4] * Ji[4] + Ji[7] * Ji[7]);
                h[4] = ww * (Ji[1] * Ji[2] + Ji[4] * Ji[5] + Ji[7] * Ji[8]);
                h[5] = ww * (Ji[2] * Ji[2] + Ji[5] * Ji[5] + Ji[8] * Ji[8]);
                h[6] = -ww * (Ji[0] * err[0] + Ji[3] * err[1] + Ji[6] * err[2]);
                h[7] = -ww * (Ji[1] * err[0] + Ji[4] * err[1] + Ji[7] * err[2]);
                h[8] = -ww * (Ji[2] * err[0] + Ji[5] * err[1] + Ji[8] * err[2]);
#pragma unroll
                for (int i = 0; i < 9; ++i) pl[(size_t)i * v.E + e] = lfree ? h[i] : 0.0;
                // Hpl = Jj^T W Ji (6x3)
#pragma unroll
                for (int a = 0; a < 6; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        const double s = ww * (Jj[a] * Ji[b] + Jj[6 + a] * Ji[3 + b] + Jj[12 + a] * Ji[6 + b]);
                        Hpl[(size_t)e * 18 + a * 3 + b] = (lfree && pfree) ? s : 0.0;
                    }
            }
        } else if (linearize) {
#pragma unroll
            for (int i = 0; i < 9; ++i) pl[(size_t)i * v.E + e] = 0.0;
#pragma unroll
            for (int i = 0; i < 18; ++i) Hpl[(size_t)e * 18 + i] = 0.0;
        } else if (chi_carry) {
            chi[e] = chi_carry[e];  // inactive edges keep the chi2 of their last activation across the current/trial swap
        }
    }
    const double s = block_sum(cost, sh);
    if (threadIdx.x == 0) chi_partials[blockIdx.x] = s;
    if (tail.ticket) {  // trial evaluation: the last CTA runs the accept / reject bookkeeping on the complete partial sums
        __shared__ int last_flag;
        __shared__ double stage[1024];
        if (!last_cta_arrives(tail.ticket, &last_flag)) return;
        lm_after_trial(tail.ctl, tail.r_chi, tail.eb, tail.r_scale, tail.lb2, tail.r_result, tail.n_launches, tail.h_inner, tail.h_outer, tail.use_graph,
                       stage);
    }
}

// K2: per landmark -- Hll (6 unique), bl (3) from its contiguous edge range; partial max |diag|
__global__ void __launch_bounds__(128) points_kernel(View v, const double* __restrict__ pl, double* __restrict__ Hll, double* __restrict__ bl,
                                                     double* __restrict__ diag_partials) {
    __shared__ double sh[128];
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    double mx = 0.0;
    if (l < v.L) {
        const int lc = v.edges[v.pt_start[l] < v.E ? v.pt_start[l] : 0].lcol;  // same for all edges of the landmark
        const int a = v.pt_start[l], b = v.pt_start[l + 1];
        if (a < b && lc >= 0) {
            double h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int e = a; e < b; ++e)
#pragma unroll
                for (int i = 0; i < 9; ++i) h[i] += pl[(size_t)i * v.E + e];
#pragma unroll
            for (int i = 0; i < 6; ++i) Hll[(size_t)i * v.Lf + lc] = h[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) bl[(size_t)i * v.Lf + lc] = h[6 + i];
            mx = fmax(fabs(h[0]), fmax(fabs(h[3]), fabs(h[5])));
        }
    }
    // max is order independent
    sh[threadIdx.x] = mx;
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) diag_partials[blockIdx.x] = sh[0];
}

// K3: keyframe-side blocks.  The edges of every free keyframe are cut into chunks of kPoseChunk; one warp reduces one
//     chunk to 21 unique Hpp entries + 6 bp entries; the chunk partials are then added in index order.
constexpr int kPoseChunk = 64;
__device__ __forceinline__ void pose_finish_element(int idx, int Kf, const int* __restrict__ chunk_start, const double* __restrict__ partials,
                                                    double* __restrict__ Hpp, double* __restrict__ bp) {
    const int p = idx / 27, t = idx - p * 27;
    if (p >= Kf) return;
    double r = 0.0;
    for (int c = chunk_start[p]; c < chunk_start[p + 1]; ++c) r += __ldcg(partials + (size_t)c * 27 + t);
    if (t < 21) {
        int a = 0, rem = t;
        while (rem >= 6 - a) {
            rem -= 6 - a;
            ++a;
        }
        const int b = a + rem;
        Hpp[(size_t)p * 36 + a * 6 + b] = r;
        Hpp[(size_t)p * 36 + b * 6 + a] = r;
    } else {
        bp[(size_t)p * 6 + (t - 21)] = r;
    }
}

// One warp per chunk; the last CTA to finish adds the chunk partials of every keyframe in index order (what used to be a second
// launch) and then runs the after-build LM bookkeeping (a third).
__global__ void __launch_bounds__(128) pose_chunks_kernel(View v, const int2* __restrict__ chunks, int n_chunks, Dual Rt2, Dual pts2,
                                                          double* __restrict__ partials, const LmCtl* __restrict__ ctl,
                                                          const int* __restrict__ chunk_start, double* __restrict__ Hpp, double* __restrict__ bp,
                                                          CtlTail tail) {
    const double* __restrict__ Rt = Rt2.p[ctl->cur & 1];
    const double* __restrict__ pts = pts2.p[ctl->cur & 1];
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid < n_chunks) {
        const int2 ch = chunks[wid];
        double acc[27];
#pragma unroll
        for (int i = 0; i < 27; ++i) acc[i] = 0.0;
        for (int k = ch.x + lane; k < ch.y; k += 32) {
            const int e = v.pose_edges[k];
            if (v.level[e]) continue;
            const EdgeS ed = v.edges[e];
            const Cam c = v.cams[ed.cam];
            double err[3], pc[3], Ji[9], Jj[18];
            const double* T = Rt + 12 * (size_t)ed.pose;
            edge_residual(ed, c, T, pts + 3 * (size_t)ed.point, err, pc);
            edge_jacobians(ed, c, T, pc, Ji, Jj);
            const double w = (double)ed.inv_sigma_sq;
            const double e2 = w * (err[0] * err[0] + err[1] * err[1] + err[2] * err[2]);
            const double ww = w * (v.robust[e] ? huber_weight(e2, (double)ed.delta) : 1.0);
            int t = 0;
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int b = a; b < 6; ++b) acc[t++] += ww * (Jj[a] * Jj[b] + Jj[6 + a] * Jj[6 + b] + Jj[12 + a] * Jj[12 + b]);
#pragma unroll
            for (int a = 0; a < 6; ++a) acc[21 + a] += -ww * (Jj[a] * err[0] + Jj[6 + a] * err[1] + Jj[12 + a] * err[2]);
        }
#pragma unroll
        for (int i = 0; i < 27; ++i)
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) acc[i] += __shfl_down_sync(0xFFFFFFFFu, acc[i], s);
        if (lane == 0)
#pragma unroll
            for (int i = 0; i < 27; ++i) partials[(size_t)wid * 27 + i] = acc[i];
    }
    __shared__ int last_flag;
    __shared__ double stage[1024];
    if (!last_cta_arrives(tail.ticket, &last_flag)) return;
    for (int idx = threadIdx.x; idx < tail.Kf * 27; idx += blockDim.x) pose_finish_element(idx, tail.Kf, chunk_start, partials, Hpp, bp);
    __threadfence();
    __syncthreads();
    lm_after_build(tail.ctl, tail.r_chi, tail.eb, tail.r_diag, tail.lb, tail.Hpp, tail.Kf, tail.n_launches, tail.h_inner, tail.use_graph, stage);
}

// K4: per landmark -- Dinv = (Hll + lambda I)^-1 (symmetric 3x3, cofactor inverse like Eigen's fixed-size path)
__global__ void __launch_bounds__(128) dinv_kernel(int Lf, const LmCtl* __restrict__ ctl, const double* __restrict__ Hll,
                                                   double* __restrict__ Dinv, int* __restrict__ fail) {
    const int lc = blockIdx.x * blockDim.x + threadIdx.x;
    if (lc >= Lf) return;
    const double lambda = ctl->lambda;
    const double A0 = Hll[lc] + lambda, A1 = Hll[(size_t)Lf + lc], A2 = Hll[(size_t)2 * Lf + lc];
    const double A4 = Hll[(size_t)3 * Lf + lc] + lambda, A5 = Hll[(size_t)4 * Lf + lc], A8 = Hll[(size_t)5 * Lf + lc] + lambda;
    const double c0 = A4 * A8 - A5 * A5, c1 = A5 * A2 - A1 * A8, c2 = A1 * A5 - A4 * A2;
    const double det = A0 * c0 + A1 * c1 + A2 * c2;
    if (det == 0.0 || !isfinite(det)) {
        *fail = 1;
        return;
    }
    const double id = 1.0 / det;
    Dinv[lc] = c0 * id;
    Dinv[(size_t)Lf + lc] = c1 * id;
    Dinv[(size_t)2 * Lf + lc] = c2 * id;
    Dinv[(size_t)3 * Lf + lc] = (A0 * A8 - A2 * A2) * id;
    Dinv[(size_t)4 * Lf + lc] = (A1 * A2 - A0 * A5) * id;
    Dinv[(size_t)5 * Lf + lc] = (A0 * A4 - A1 * A1) * id;
}

// K5: Schur complement of the landmarks.  The (edge a, edge c) pairs that share a landmark are grouped by the upper
//     block (i <= j) of the reduced system they fall into and cut into chunks of kSchurChunk pairs; one warp reduces one
//     chunk:  partial = sum T(a) Hpl(c)^T,  T(a) = Hpl(a) Dinv(l)   (+ for diagonal blocks  sum T(a) bl(l)).
//     The chunk partials of every block are then added in index order (deterministic) into
//     M = [Hpp + lambda I - sum ; (bp - sum)^T].
constexpr int kSchurChunk = 64;
struct SchurBlock {
    int i, j, chunk_start, chunk_end;
};
struct SchurChunk {
    int start, end, diag, pad;
};
__device__ __forceinline__ void load18(const double* __restrict__ p, double* out) {
    const double2* q = reinterpret_cast<const double2*>(p);  // 144-byte records, 16-byte aligned
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const double2 v = q[i];
        out[2 * i] = v.x;
        out[2 * i + 1] = v.y;
    }
}
// Adds the chunk partials of block b in index order into M = [Hpp + lambda I - sum ; (bp - sum)^T]; lanes cover the 42 elements.
__device__ __forceinline__ void schur_finish_block(const SchurBlock sb, double lambda, const double* __restrict__ partials,
                                                   const double* __restrict__ Hpp, const double* __restrict__ bp, double* __restrict__ M, int n, int ld,
                                                   int lane) {
    for (int el = lane; el < 42; el += 32) {
        if (el >= 36 && sb.i != sb.j) continue;
        double sacc = 0.0;
        for (int c = sb.chunk_start; c < sb.chunk_end; ++c) sacc += __ldcg(partials + (size_t)c * 42 + el);
        if (el < 36) {
            const int r = el / 6, c = el - r * 6;
            double val = -sacc;
            if (sb.i == sb.j) val += Hpp[(size_t)sb.i * 36 + el] + (r == c ? lambda : 0.0);
            M[(size_t)(6 * sb.j + c) * ld + 6 * sb.i + r] = val;  // lower triangle (j >= i)
            if (sb.i == sb.j) M[(size_t)(6 * sb.i + r) * ld + 6 * sb.j + c] = val;
        } else {
            const int r = el - 36;
            M[(size_t)n * ld + 6 * sb.i + r] = bp[(size_t)sb.i * 6 + r] - sacc;  // rhs row
        }
    }
}

// One warp per chunk (chunk.pad = its block).  The warp that completes a block's last chunk assembles that block of the reduced
// system (what used to be a second launch); warps past the chunks fill the blocks that have no pair at all.
__global__ void __launch_bounds__(128) schur_chunks_kernel(View v, const SchurChunk* __restrict__ chunks, int n_chunks,
                                                           const int2* __restrict__ pairs, const double* __restrict__ Hpl,
                                                           const double* __restrict__ Dinv, const double* __restrict__ bl,
                                                           double* __restrict__ partials, const LmCtl* __restrict__ ctl,
                                                           const SchurBlock* __restrict__ blocks, const int* __restrict__ empty_blocks, int n_empty,
                                                           int* __restrict__ blk_done, const double* __restrict__ Hpp, const double* __restrict__ bp,
                                                           double* __restrict__ M, int n, int ld) {
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= n_chunks) {
        if (wid - n_chunks < n_empty) schur_finish_block(blocks[empty_blocks[wid - n_chunks]], ctl->lambda, partials, Hpp, bp, M, n, ld, lane);
        return;
    }
    const SchurChunk ch = chunks[wid];
    double acc[42];
#pragma unroll
    for (int i = 0; i < 42; ++i) acc[i] = 0.0;
    for (int k = ch.start + lane; k < ch.end; k += 32) {
        const int2 pr = pairs[k];
        double ha[18], hc[18], t[18];
        load18(Hpl + (size_t)pr.x * 18, ha);
        load18(Hpl + (size_t)pr.y * 18, hc);
        const int lc = v.edges[pr.x].lcol;
        const double D0 = Dinv[lc], D1 = Dinv[(size_t)v.Lf + lc], D2 = Dinv[(size_t)2 * v.Lf + lc];
        const double D4 = Dinv[(size_t)3 * v.Lf + lc], D5 = Dinv[(size_t)4 * v.Lf + lc], D8 = Dinv[(size_t)5 * v.Lf + lc];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            t[r * 3] = ha[r * 3] * D0 + ha[r * 3 + 1] * D1 + ha[r * 3 + 2] * D2;
            t[r * 3 + 1] = ha[r * 3] * D1 + ha[r * 3 + 1] * D4 + ha[r * 3 + 2] * D5;
            t[r * 3 + 2] = ha[r * 3] * D2 + ha[r * 3 + 1] * D5 + ha[r * 3 + 2] * D8;
        }
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = 0; c < 6; ++c) acc[r * 6 + c] += t[r * 3] * hc[c * 3] + t[r * 3 + 1] * hc[c * 3 + 1] + t[r * 3 + 2] * hc[c * 3 + 2];
        if (ch.diag) {  // pr.x == pr.y: the edge of keyframe i to this landmark
            const double b0 = bl[lc], b1 = bl[(size_t)v.Lf + lc], b2 = bl[(size_t)2 * v.Lf + lc];
#pragma unroll
            for (int r = 0; r < 6; ++r) acc[36 + r] += t[r * 3] * b0 + t[r * 3 + 1] * b1 + t[r * 3 + 2] * b2;
        }
    }
#pragma unroll
    for (int i = 0; i < 42; ++i)
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc[i] += __shfl_down_sync(0xFFFFFFFFu, acc[i], s);
    const SchurBlock sb = blocks[ch.pad];
    int arrived = 0;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 42; ++i) partials[(size_t)wid * 42 + i] = acc[i];
        __threadfence();
        arrived = atomicAdd(&blk_done[ch.pad], 1);
    }
    arrived = __shfl_sync(0xFFFFFFFFu, arrived, 0);
    if (arrived != sb.chunk_end - sb.chunk_start - 1) return;
    __threadfence();
    schur_finish_block(sb, ctl->lambda, partials, Hpp, bp, M, n, ld, lane);
    if (lane == 0) blk_done[ch.pad] = 0;  // re-armed for the next trial
}

// K6: dense Cholesky of the reduced system (<= 6*Kf unknowns), solve, then the keyframe updates
//     (shot_vertex::oplusImpl: T <- exp(dx) * T) into the trial state.  One CTA.
__device__ void se3_oplus(const double* q, const double* t, const double* upd, double* qo, double* to) {
    const double* om = upd;
    const double* up = upd + 3;
    const double theta = sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
    const double O[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
    double O2[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) O2[i * 3 + j] = O[i * 3] * O[j] + O[i * 3 + 1] * O[3 + j] + O[i * 3 + 2] * O[6 + j];
    double a, b, c, d;
    if (theta < 0.00001) {  // g2o SE3Quat::exp small-angle branch
        a = 1.0; b = 0.5; c = 0.5; d = 1.0 / 6.0;
    } else {
        a = sin(theta) / theta;
        b = (1 - cos(theta)) / (theta * theta);
        c = b;
        d = (theta - sin(theta)) / (theta * theta * theta);
    }
    double R[9], V[9];
    for (int i = 0; i < 9; ++i) {
        const double I = (i % 4 == 0) ? 1.0 : 0.0;
        R[i] = I + a * O[i] + b * O2[i];
        V[i] = I + c * O[i] + d * O2[i];
    }
    double dq[4], dt[3];
    rot_to_quat(R, dq);
    quat_normalize(dq);
    for (int i = 0; i < 3; ++i) dt[i] = V[i * 3] * up[0] + V[i * 3 + 1] * up[1] + V[i * 3 + 2] * up[2];
    // (dq, dt) * (q, t): rotate t by dq (Eigen: v + w*uv + qv x uv, uv = 2 qv x v)
    double uv[3] = {dq[1] * t[2] - dq[2] * t[1], dq[2] * t[0] - dq[0] * t[2], dq[0] * t[1] - dq[1] * t[0]};
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    to[0] = dt[0] + t[0] + dq[3] * uv[0] + (dq[1] * uv[2] - dq[2] * uv[1]);
    to[1] = dt[1] + t[1] + dq[3] * uv[1] + (dq[2] * uv[0] - dq[0] * uv[2]);
    to[2] = dt[2] + t[2] + dq[3] * uv[2] + (dq[0] * uv[1] - dq[1] * uv[0]);
    double nq[4];
    nq[3] = dq[3] * q[3] - dq[0] * q[0] - dq[1] * q[1] - dq[2] * q[2];
    nq[0] = dq[3] * q[0] + dq[0] * q[3] + dq[1] * q[2] - dq[2] * q[1];
    nq[1] = dq[3] * q[1] + dq[1] * q[3] + dq[2] * q[0] - dq[0] * q[2];
    nq[2] = dq[3] * q[2] + dq[2] * q[3] + dq[0] * q[1] - dq[1] * q[0];
    quat_normalize(nq);
    qo[0] = nq[0]; qo[1] = nq[1]; qo[2] = nq[2]; qo[3] = nq[3];
}

constexpr int kCholThreads = 512;
constexpr int kNB = 24;       // panel width: four 6x6 keyframe blocks
constexpr int kCholCluster = 8;  // CTAs (SMs) that share one factorisation

__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_nctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
// all threads of all CTAs of the cluster; release/acquire at cluster scope orders the global-memory updates of the trailing matrix
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Blocked right-looking Cholesky of the augmented matrix M = [Hs ; bs^T] ((n+1) x ld, row-major, lower triangle): the
// right-hand side rides along as row n, so after the factorisation M[n][0..n) = L^-1 bs and only the backward solve
// remains.  Per panel of kNB columns: (1) one warp factors the diagonal block in registers (lane i owns row i, columns
// are exchanged by shuffle); (2) every row below is solved against it, column oriented, with the reciprocal diagonal;
// the solved panel is kept TRANSPOSED in shared memory; (3) rank-kNB trailing update with 4x4 register tiles whose
// operands are two 32-byte vector loads per panel column.
// The kernel runs as ONE thread-block cluster: every CTA repeats the cheap steps (1) and (2) on its own SM (so no panel
// exchange is needed), the tiles of step (3) -- 60 % of the flops -- are dealt round-robin to the CTAs of the cluster, and a
// cluster barrier (release/acquire) separates the panels.  CTA 0 writes the factor back and does the backward solve.
__global__ void __launch_bounds__(kCholThreads) chol_solve_kernel(int n, int ld, double* __restrict__ M, const LmCtl* __restrict__ ctl,
                                                                  const double* __restrict__ bp, double* __restrict__ xp, int K,
                                                                  const int* __restrict__ pose_col, Dual q2, Dual t2, Dual Rt2,
                                                                  double* __restrict__ result, int* __restrict__ fail) {
    extern __shared__ __align__(32) double dyn[];
    const double lambda = ctl->lambda;
    const int cur_idx = ctl->cur & 1;
    const double* __restrict__ q_cur = q2.p[cur_idx];
    const double* __restrict__ t_cur = t2.p[cur_idx];
    double* __restrict__ q_new = q2.p[cur_idx ^ 1];
    double* __restrict__ t_new = t2.p[cur_idx ^ 1];
    double* __restrict__ Rt_new = Rt2.p[cur_idx ^ 1];
    double* D = dyn;                     // kNB x (kNB+1) diagonal block (lower, padded with the identity)
    double* Pn = dyn + kNB * (kNB + 1);  // kNB x mp panel, transposed (k-major); the 600 doubles in front keep it 32-byte aligned
    const int mp = (n + 1 + 3) & ~3;
    __shared__ double xs[1024];
    __shared__ double invd_all[1024];    // 1 / L[j][j]
    __shared__ double sh[kCholThreads];
    __shared__ int bad;
    const int tid = threadIdx.x, nt = blockDim.x;
    const unsigned crank = cluster_ctarank(), csize = cluster_nctarank();
    const bool lead = crank == 0;
    if (tid == 0) bad = *fail;
    __syncthreads();
    long long t_diag = 0, t_panel = 0, t_trail = 0, t_back = 0, t0 = clock64(), t1;
#define PHASE(acc) do { t1 = clock64(); acc += t1 - t0; t0 = t1; } while (0)
    for (int kb = 0; kb < n && !bad; kb += kNB) {
        const int nb = min(kNB, n - kb);
        if (tid < 32) {
            // (1) diagonal block, warp 0.  A short last block is padded with the identity so everything is fully unrolled.
            double r[kNB];
#pragma unroll
            for (int k = 0; k < kNB; ++k) r[k] = (tid < nb && k <= tid) ? M[(size_t)(kb + tid) * ld + kb + k] : ((k == tid) ? 1.0 : 0.0);
            int b = 0;
#pragma unroll
            for (int j = 0; j < kNB; ++j) {
                const double djj = __shfl_sync(0xFFFFFFFFu, r[j], j);
                b |= (!(djj > 0.0) || !isfinite(djj)) ? 1 : 0;
                const double inv = rsqrt(djj), dd = djj * inv;  // one reciprocal square root instead of sqrt + divide
                if (tid == 0) invd_all[kb + j] = inv;  // (entries past n are never read)
                r[j] = (tid == j) ? dd : ((tid > j) ? r[j] * inv : r[j]);
                const double mine = (tid > j) ? r[j] : 0.0;
#pragma unroll
                for (int k = j + 1; k < kNB; ++k) {
                    const double lkj = __shfl_sync(0xFFFFFFFFu, r[j], k);
                    r[k] = fma(-((tid >= k) ? mine : 0.0), lkj, r[k]);
                }
            }
            if (tid < kNB) {
#pragma unroll
                for (int k = 0; k < kNB; ++k) {
                    D[tid * (kNB + 1) + k] = r[k];
                }
            }
            if (tid == 0 && b) bad = 1;
        }
        __syncthreads();
        PHASE(t_diag);
        if (bad) break;
        // (2) panel solve: every row below the block (including the rhs row n)
        const int r0 = kb + nb, m = n + 1 - r0;
        for (int t = tid; t < m; t += nt) {
            double* row = M + (size_t)(r0 + t) * ld + kb;
            double x[kNB];
#pragma unroll
            for (int j = 0; j < kNB; ++j) x[j] = (j < nb) ? row[j] : 0.0;
#pragma unroll
            for (int j = 0; j < kNB; ++j) {
                x[j] *= invd_all[kb + j];
#pragma unroll
                for (int k = j + 1; k < kNB; ++k) x[k] = fma(-x[j], D[k * (kNB + 1) + j], x[k]);
            }
#pragma unroll
            for (int j = 0; j < kNB; ++j) {
                Pn[(size_t)j * mp + t] = x[j];
            }
        }
        for (int idx = tid; idx < kNB * 4; idx += nt) {  // zero the <= 3 padding rows read by the last 4-row tile
            const int j = idx >> 2, t = m + (idx & 3);
            if (t < mp) Pn[(size_t)j * mp + t] = 0.0;
        }
        __syncthreads();
        PHASE(t_panel);
        // (3) trailing update with 4x4 register tiles over the lower triangle (rhs row included, rhs column excluded)
        const int tm = (m + 3) >> 2;
        const int n_tiles = tm * (tm + 1) / 2;
        for (int tile = tid * csize + crank; tile < n_tiles; tile += nt * csize) {
            int tr = (int)((sqrt(8.0 * tile + 1.0) - 1.0) * 0.5);
            while ((tr + 1) * (tr + 2) / 2 <= tile) ++tr;
            while (tr * (tr + 1) / 2 > tile) --tr;
            const int tc = tile - tr * (tr + 1) / 2;
            double acc[16];
#pragma unroll
            for (int a = 0; a < 16; ++a) acc[a] = 0.0;
            const int rb = tr * 4, cb = tc * 4;
#pragma unroll 4
            for (int k = 0; k < kNB; ++k) {
                const double2* pa = reinterpret_cast<const double2*>(Pn + (size_t)k * mp + rb);
                const double2* pb = reinterpret_cast<const double2*>(Pn + (size_t)k * mp + cb);
                const double2 a0 = pa[0], a1 = pa[1], b0 = pb[0], b1 = pb[1];
                const double a[4] = {a0.x, a0.y, a1.x, a1.y}, b[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int w = 0; w < 4; ++w) acc[u * 4 + w] = fma(a[u], b[w], acc[u * 4 + w]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const int r = rb + u, c = cb + w;
                    if (r < m && c <= r && r0 + c < n) M[(size_t)(r0 + r) * ld + r0 + c] -= acc[u * 4 + w];
                }
        }
        __syncthreads();
        if (csize > 1) cluster_sync_all();  // every CTA's share of the trailing matrix is visible before the next panel is read
        // The factor of this panel is written back only now: the other CTAs were still reading these columns of M (their own copy
        // of steps 1-2) until the barrier; the next panel touches columns >= kb + nb only.
        if (lead) {
            for (int idx = tid; idx < nb * nb; idx += nt) {
                const int i = idx / nb, j = idx - i * nb;
                if (j <= i) M[(size_t)(kb + i) * ld + kb + j] = D[i * (kNB + 1) + j];
            }
            for (int idx = tid; idx < m * nb; idx += nt) {
                const int t = idx / nb, j = idx - t * nb;
                M[(size_t)(r0 + t) * ld + kb + j] = Pn[(size_t)j * mp + t];
            }
        }
        __syncthreads();
        PHASE(t_trail);
    }
    if (!lead) return;  // the backward solve and the keyframe updates are one CTA's work
    if (bad) {
        if (tid == 0) {
            *fail = 1;
            result[0] = 0.0;
        }
    } else {
        // backward solve L^T x = y, y = M[n][0..n)
        for (int i = tid; i < n; i += nt) xs[i] = M[(size_t)n * ld + i];
        __syncthreads();
        for (int kb = ((n - 1) / kNB) * kNB; kb >= 0; kb -= kNB) {
            const int nb = min(kNB, n - kb);
            for (int idx = tid; idx < kNB * kNB; idx += nt) {
                const int i = idx / kNB, j = idx - i * kNB;
                D[i * (kNB + 1) + j] = (i < nb && j <= i) ? M[(size_t)(kb + i) * ld + kb + j] : 0.0;
            }
            __syncthreads();
            if (tid < 32) {  // diagonal block: lane k owns y[k]; x[j] is broadcast by shuffle
                double y = (tid < nb) ? xs[kb + tid] : 0.0;
#pragma unroll
                for (int j = kNB - 1; j >= 0; --j) {
                    const double xj = __shfl_sync(0xFFFFFFFFu, y, j) * ((j < nb) ? invd_all[kb + j] : 0.0);
                    if (tid == j) y = xj;
                    else if (tid < j) y = fma(-D[j * (kNB + 1) + tid], xj, y);
                }
                if (tid < nb) xs[kb + tid] = y;
            }
            __syncthreads();
            for (int i = tid; i < kb; i += nt) {
                double sacc = xs[i];
                for (int k = 0; k < nb; ++k) sacc = fma(-M[(size_t)(kb + k) * ld + i], xs[kb + k], sacc);
                xs[i] = sacc;
            }
            __syncthreads();
        }
        double sc = 0.0;  // pose part of computeScale: sum x (lambda x + b)
        for (int i = tid; i < n; i += nt) {
            xp[i] = xs[i];
            sc += xs[i] * (lambda * xs[i] + bp[i]);
        }
        const double tot = block_sum(sc, sh);
        PHASE(t_back);
        if (tid == 0) {
            result[0] = 1.0;
            result[1] = tot;
            result[2] = (double)t_diag;
            result[3] = (double)t_panel;
            result[4] = (double)t_trail;
            result[5] = (double)t_back;
        }
    }
#undef PHASE
    __syncthreads();
    // trial keyframe states (fixed keyframes and failed solves keep the current state)
    for (int k = tid; k < K; k += nt) {
        double qn[4], tn[3];
        const int pc = pose_col[k];
        if (pc >= 0 && !bad) {
            se3_oplus(q_cur + 4 * k, t_cur + 3 * k, xs + 6 * pc, qn, tn);
        } else {
            for (int i = 0; i < 4; ++i) qn[i] = q_cur[4 * k + i];
            for (int i = 0; i < 3; ++i) tn[i] = t_cur[3 * k + i];
        }
        for (int i = 0; i < 4; ++i) q_new[4 * k + i] = qn[i];
        for (int i = 0; i < 3; ++i) t_new[3 * k + i] = tn[i];
        double R[9];
        quat_to_rot(qn, R);
        for (int i = 0; i < 9; ++i) Rt_new[12 * k + i] = R[i];
        for (int i = 0; i < 3; ++i) Rt_new[12 * k + 9 + i] = tn[i];
    }
}

// K7: back-substitution x_l = Dinv (bl - sum_e Hpl(e)^T x_p), trial landmark, scale partials.  Eight lanes share a landmark
//     (they split its edges), sixteen landmarks per 128-thread CTA.
__global__ void __launch_bounds__(128) backsub_kernel(View v, const LmCtl* __restrict__ ctl, const double* __restrict__ Dinv,
                                                      const double* __restrict__ bl, const double* __restrict__ Hpl,
                                                      const double* __restrict__ xp, Dual pts2, double* __restrict__ scale_partials,
                                                      const int* __restrict__ fail) {
    __shared__ double sh[128];
    const double lambda = ctl->lambda;
    const double* __restrict__ pts_cur = pts2.p[ctl->cur & 1];
    double* __restrict__ pts_new = pts2.p[(ctl->cur & 1) ^ 1];
    const int sub = threadIdx.x & 7;
    const int l = blockIdx.x * 16 + (threadIdx.x >> 3);
    double sc = 0.0;
    const bool valid = l < v.L;
    const int a0 = valid ? v.pt_start[l] : 0, b0 = valid ? v.pt_start[l + 1] : 0;
    const int lc = (a0 < b0) ? v.edges[a0].lcol : -1;
    const bool solve = lc >= 0 && !*fail;
    double c0 = 0.0, c1 = 0.0, c2 = 0.0;
    if (solve) {
        for (int e = a0 + sub; e < b0; e += 8) {
            const int pcol = v.edges[e].pcol;
            if (pcol < 0) continue;
            double h[18];
            load18(Hpl + (size_t)e * 18, h);
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                const double x = xp[6 * pcol + r];
                c0 -= h[r * 3] * x;
                c1 -= h[r * 3 + 1] * x;
                c2 -= h[r * 3 + 2] * x;
            }
        }
    }
    // fixed-order reduction over the 8 lanes of the landmark
#pragma unroll
    for (int s = 4; s > 0; s >>= 1) {
        c0 += __shfl_down_sync(0xFFFFFFFFu, c0, s, 8);
        c1 += __shfl_down_sync(0xFFFFFFFFu, c1, s, 8);
        c2 += __shfl_down_sync(0xFFFFFFFFu, c2, s, 8);
    }
    if (valid && sub == 0) {
        double p0 = pts_cur[3 * (size_t)l], p1 = pts_cur[3 * (size_t)l + 1], p2 = pts_cur[3 * (size_t)l + 2];
        if (solve) {
            const double bb0 = bl[lc], bb1 = bl[(size_t)v.Lf + lc], bb2 = bl[(size_t)2 * v.Lf + lc];
            c0 += bb0; c1 += bb1; c2 += bb2;
            const double D0 = Dinv[lc], D1 = Dinv[(size_t)v.Lf + lc], D2 = Dinv[(size_t)2 * v.Lf + lc];
            const double D4 = Dinv[(size_t)3 * v.Lf + lc], D5 = Dinv[(size_t)4 * v.Lf + lc], D8 = Dinv[(size_t)5 * v.Lf + lc];
            const double x0 = D0 * c0 + D1 * c1 + D2 * c2, x1 = D1 * c0 + D4 * c1 + D5 * c2, x2 = D2 * c0 + D5 * c1 + D8 * c2;
            sc = x0 * (lambda * x0 + bb0) + x1 * (lambda * x1 + bb1) + x2 * (lambda * x2 + bb2);
            p0 += x0; p1 += x1; p2 += x2;  // landmark_vertex::oplusImpl
        }
        pts_new[3 * (size_t)l] = p0; pts_new[3 * (size_t)l + 1] = p1; pts_new[3 * (size_t)l + 2] = p2;
    }
    const double tot = block_sum(sc, sh);
    if (threadIdx.x == 0) scale_partials[blockIdx.x] = tot;
}

// K8: outlier test (local_bundle_adjuster_g2o.cc:323-344, 357-375): chi2 of the last activation vs the chi-square
//     threshold, or non-positive depth at the current estimate.  mode 0: mark level + drop the kernel; mode 1: report.
__global__ void __launch_bounds__(128) outlier_kernel(View v, Dual Rt2, Dual pts2, Dual chi2, int mode, unsigned char* __restrict__ out,
                                                      const LmCtl* __restrict__ ctl) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= v.E) return;
    if (mode == 0 && ctl->skip_round2) return;  // local_bundle_adjuster_g2o.cc:317-321: no second round after an abort
    const double* __restrict__ Rt = Rt2.p[ctl->cur & 1];
    const double* __restrict__ pts = pts2.p[ctl->cur & 1];
    const double* __restrict__ chi = chi2.p[ctl->cur & 1];
    const EdgeS ed = v.edges[e];
    unsigned char o = 0;
    if (ed.can_outlier) {
        const double thr = (ed.oxr < 0.f) ? (double)5.99146f : (double)7.81473f;
        bool depth_ok = true;
        if (v.cams[ed.cam].model != 1) {  // reproj_edge_wrapper.h:233-268 (equirectangular: always true)
            const double* T = Rt + 12 * (size_t)ed.pose;
            const double* P = pts + 3 * (size_t)ed.point;
            depth_ok = 0.0 < T[6] * P[0] + T[7] * P[1] + T[8] * P[2] + T[11];
        }
        o = (thr < chi[e] || !depth_ok) ? 1 : 0;
    }
    if (mode == 0) {
        if (ed.can_outlier) {
            if (o) v.level[e] = 1;
            v.robust[e] = 0;
        }
    } else {
        out[e] = o;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// optimize::pose_optimizer: motion-only BA of one frame, the step between the two per-frame matcher calls.
//   pose_optimizer_g2o::optimize            src/stella_vslam/optimize/pose_optimizer_g2o.cc:38-175
//   mono / stereo_perspective_pose_opt_edge optimize/internal/se3/perspective_pose_opt_edge.h  (= pose block of the reprojection edges)
//   equirectangular_pose_opt_edge           optimize/internal/se3/equirectangular_pose_opt_edge.h
// Six unknowns: the whole protocol -- (num_trials_robust + num_trials) calls of optimize(num_each_iter) with LM, the terminate
// action and the outlier re-classification in between -- runs inside ONE kernel launch, one CTA per frame, no host round trip.
// Sums over edges are per-thread strided partials combined in a fixed order (deterministic).
// ---------------------------------------------------------------------------------------------------------------
struct PoseEdge {
    double pw[3];
    float ox, oy, oxr, inv_sigma_sq, delta;
    int pad;
};
struct PoseProb {
    int n;
    int edge_off;  // into the flat edge / flag arrays
    Cam cam;
    double q[4], t[3];
};
constexpr int kPoseThreads = 256;

__device__ __forceinline__ void pose_rt(const double* q, const double* t, double* Rt) {
    quat_to_rot(q, Rt);
    Rt[9] = t[0];
    Rt[10] = t[1];
    Rt[11] = t[2];
}
// sum of v over the CTA in a fixed order: warp shuffle tree, then the warp leaders in index order
template <int N>
__device__ __forceinline__ void cta_sum(double (&v)[N], double* out, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int s2 = 16; s2 > 0; s2 >>= 1) v[i] += __shfl_down_sync(0xFFFFFFFFu, v[i], s2);
    }
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < N; ++i) scratch[warp * N + i] = v[i];
    __syncthreads();
    if (threadIdx.x < N) {
        double r = 0.0;
        for (int w = 0; w < kPoseThreads / 32; ++w) r += scratch[w * N + threadIdx.x];
        out[threadIdx.x] = r;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kPoseThreads) pose_optimize_kernel(const PoseProb* __restrict__ probs, const PoseEdge* __restrict__ edges_all,
                                                                     unsigned char* __restrict__ level_all, unsigned char* __restrict__ flags_all,
                                                                     int trials_robust, int trials, int each_iter, double* __restrict__ pose_out,
                                                                     unsigned* __restrict__ n_valid_out) {
    __shared__ double scratch[(kPoseThreads / 32) * 28];
    __shared__ double red[28];
    __shared__ double Rt[12], Rt_trial[12], q_cur[4], t_cur[3], q_trial[4], t_trial[3], x[6];
    __shared__ double s_lambda, s_ni, s_cur_chi, s_last_chi, s_rho;
    __shared__ int s_ok2, s_go_inner, s_go_outer, s_accept, s_it, s_qmax, s_ok, s_stop, s_bad;
    const PoseProb pb = probs[blockIdx.x];
    const PoseEdge* __restrict__ edges = edges_all + pb.edge_off;
    unsigned char* __restrict__ level = level_all + pb.edge_off;
    unsigned char* __restrict__ flags = flags_all + pb.edge_off;
    const int tid = threadIdx.x, n = pb.n;
    const Cam cam = pb.cam;
    if (tid < 4) q_cur[tid] = pb.q[tid];
    if (tid < 3) t_cur[tid] = pb.t[tid];
    for (int e = tid; e < n; e += kPoseThreads) {
        level[e] = 0;
        flags[e] = 0;
    }
    __syncthreads();
    if (n < 5) {  // pose_optimizer_g2o.cc:116-118
        if (tid == 0) n_valid_out[blockIdx.x] = 0;
        if (tid < 16) {
            double Rm[9];
            quat_to_rot(pb.q, Rm);
            const int r = tid >> 2, c = tid & 3;
            pose_out[16 * (size_t)blockIdx.x + tid] = r == 3 ? (c == 3 ? 1.0 : 0.0) : (c == 3 ? pb.t[r] : Rm[r * 3 + c]);
        }
        return;
    }
    bool robust_on = trials_robust != 0;  // :123-127
    // residual / chi2 / (optionally) the normal equations of this thread's edges at pose T
    auto accumulate = [&](const double* T, bool linearize, double (&acc)[28]) {
#pragma unroll
        for (int i = 0; i < 28; ++i) acc[i] = 0.0;
        for (int e = tid; e < n; e += kPoseThreads) {
            if (level[e]) continue;
            const PoseEdge pe = edges[e];
            EdgeS ed;
            ed.ox = pe.ox; ed.oy = pe.oy; ed.oxr = pe.oxr;
            double err[3], pc[3];
            edge_residual(ed, cam, T, pe.pw, err, pc);
            const double w = (double)pe.inv_sigma_sq;
            const double e2 = w * (err[0] * err[0] + err[1] * err[1] + err[2] * err[2]);
            acc[27] += robust_on ? huber_cost(e2, (double)pe.delta) : e2;
            if (linearize) {
                double Ji[9], Jj[18];
                edge_jacobians(ed, cam, T, pc, Ji, Jj);
                const double ww = w * (robust_on ? huber_weight(e2, (double)pe.delta) : 1.0);
                int k = 0;
#pragma unroll
                for (int a = 0; a < 6; ++a)
#pragma unroll
                    for (int b = a; b < 6; ++b) acc[k++] += ww * (Jj[a] * Jj[b] + Jj[6 + a] * Jj[6 + b] + Jj[12 + a] * Jj[12 + b]);
#pragma unroll
                for (int a = 0; a < 6; ++a) acc[21 + a] += -ww * (Jj[a] * err[0] + Jj[6 + a] * err[1] + Jj[12 + a] * err[2]);
            }
        }
    };
    int bad_total = 0;
    for (int trial = 0; trial < trials_robust + trials; ++trial) {
        // ---- SparseOptimizer::optimize(each_iter) with OptimizationAlgorithmLevenberg + terminate_action
        if (tid == 0) {
            s_it = 0;
            s_ok = 1;
            s_stop = 0;
            s_go_outer = each_iter > 0;
        }
        __syncthreads();
        while (s_go_outer) {
            if (tid == 0) pose_rt(q_cur, t_cur, Rt);
            __syncthreads();
            double acc[28];
            accumulate(Rt, true, acc);
            cta_sum<28>(acc, red, scratch);
            if (tid == 0) {
                if (s_it == 0) {  // computeLambdaInit
                    s_cur_chi = red[27];
                    double mx = 0.0;
                    int k = 0;
                    for (int a = 0; a < 6; ++a) {
                        mx = fmax(mx, fabs(red[k]));
                        k += 6 - a;
                    }
                    s_lambda = 1e-5 * mx;
                    s_ni = 2.0;
                }
                s_qmax = 0;
                s_rho = 0.0;
                s_go_inner = 1;
            }
            __syncthreads();
            while (s_go_inner) {
                if (tid == 0) {  // (H + lambda I) x = b, dense Cholesky like the reduced system of the local BA
                    double A[36], b[6];
                    int k = 0;
                    for (int a = 0; a < 6; ++a)
                        for (int c = a; c < 6; ++c) {
                            A[a * 6 + c] = red[k];
                            A[c * 6 + a] = red[k];
                            ++k;
                        }
                    for (int a = 0; a < 6; ++a) {
                        A[a * 7] += s_lambda;
                        b[a] = red[21 + a];
                    }
                    int ok2 = 1;
                    for (int j = 0; j < 6 && ok2; ++j) {
                        double d = A[j * 6 + j];
                        for (int kk = 0; kk < j; ++kk) d -= A[j * 6 + kk] * A[j * 6 + kk];
                        if (!(d > 0) || !isfinite(d)) {
                            ok2 = 0;
                            break;
                        }
                        d = sqrt(d);
                        A[j * 6 + j] = d;
                        for (int i = j + 1; i < 6; ++i) {
                            double sv = A[i * 6 + j];
                            for (int kk = 0; kk < j; ++kk) sv -= A[i * 6 + kk] * A[j * 6 + kk];
                            A[i * 6 + j] = sv / d;
                        }
                    }
                    if (ok2) {
                        for (int i = 0; i < 6; ++i) {
                            double sv = b[i];
                            for (int kk = 0; kk < i; ++kk) sv -= A[i * 6 + kk] * b[kk];
                            b[i] = sv / A[i * 7];
                        }
                        for (int i = 5; i >= 0; --i) {
                            double sv = b[i];
                            for (int kk = i + 1; kk < 6; ++kk) sv -= A[kk * 6 + i] * b[kk];
                            b[i] = sv / A[i * 7];
                        }
                        for (int i = 0; i < 6; ++i) x[i] = b[i];
                        se3_oplus(q_cur, t_cur, x, q_trial, t_trial);
                    } else {
                        for (int i = 0; i < 4; ++i) q_trial[i] = q_cur[i];
                        for (int i = 0; i < 3; ++i) t_trial[i] = t_cur[i];
                    }
                    s_ok2 = ok2;
                    pose_rt(q_trial, t_trial, Rt_trial);
                }
                __syncthreads();
                double tacc[28];
                accumulate(Rt_trial, false, tacc);
                double chi1[1] = {tacc[27]};
                cta_sum<1>(chi1, red + 27, scratch);  // red[0..26] (H, b of the current state) stay valid for the next trial
                if (tid == 0) {
                    const bool ok2 = s_ok2 != 0;
                    const double temp_chi = ok2 ? red[27] : 1.7976931348623157e308;
                    double rho = s_cur_chi - temp_chi;
                    double scale = 0.0;  // computeScale
                    if (ok2)
                        for (int i = 0; i < 6; ++i) scale += x[i] * (s_lambda * x[i] + red[21 + i]);
                    scale += 1e-3;
                    rho /= scale;
                    bool broke = false;
                    if (rho > 0 && isfinite(temp_chi) && ok2) {
                        double alpha = 1. - pow(2 * rho - 1, 3.0);
                        alpha = fmin(alpha, 2. / 3.);
                        s_lambda *= fmax(1. / 3., alpha);
                        s_ni = 2.0;
                        s_cur_chi = temp_chi;
                        for (int i = 0; i < 4; ++i) q_cur[i] = q_trial[i];
                        for (int i = 0; i < 3; ++i) t_cur[i] = t_trial[i];
                    } else {
                        s_lambda *= s_ni;
                        s_ni *= 2.0;
                        if (!isfinite(s_lambda)) broke = true;
                    }
                    if (!broke) s_qmax++;
                    s_rho = rho;
                    const bool again = !broke && rho < 0 && s_qmax < 10 && !s_stop;
                    s_go_inner = again ? 1 : 0;
                    if (!again) {
                        if (s_qmax == 10 || rho == 0 || !isfinite(s_lambda)) s_ok = 0;
                        const double chi_now = s_cur_chi;
                        if (s_it == 0) {
                            s_last_chi = chi_now;
                        } else {
                            const double gain = (s_last_chi - chi_now) / chi_now;
                            s_last_chi = chi_now;
                            if (gain >= 0 && gain < 1e-3) s_stop = 1;
                        }
                        s_it++;
                        s_go_outer = (s_it < each_iter && !s_stop && s_ok) ? 1 : 0;
                    }
                }
                __syncthreads();
            }
        }
        // ---- :133-167 classify every observation at the optimised pose (inactive edges are re-evaluated, :137-139)
        if (tid == 0) {
            pose_rt(q_cur, t_cur, Rt);
            s_bad = 0;
        }
        __syncthreads();
        int bad = 0;
        for (int e = tid; e < n; e += kPoseThreads) {
            const PoseEdge pe = edges[e];
            EdgeS ed;
            ed.ox = pe.ox; ed.oy = pe.oy; ed.oxr = pe.oxr;
            double err[3], pc[3];
            edge_residual(ed, cam, Rt, pe.pw, err, pc);
            const double e2 = (double)pe.inv_sigma_sq * (err[0] * err[0] + err[1] * err[1] + err[2] * err[2]);
            const double thr = (pe.oxr < 0.f) ? (double)5.99146f : (double)7.81473f;
            const unsigned char o = thr < e2 ? 1 : 0;
            flags[e] = o;
            level[e] = o;
            bad += o;
        }
        atomicAdd(&s_bad, bad);
        if (trials != 0 && trial + 1 == trials_robust) robust_on = false;  // :164-166
        __syncthreads();
        bad_total = s_bad;
        __syncthreads();
        if (n - bad_total < 5) break;  // :169-171
    }
    if (tid == 0) {
        n_valid_out[blockIdx.x] = (unsigned)(n - bad_total);
        double Rm[9];
        quat_to_rot(q_cur, Rm);
        double* M = pose_out + 16 * (size_t)blockIdx.x;  // util::converter::to_eigen_mat
        M[0] = Rm[0]; M[1] = Rm[1]; M[2] = Rm[2]; M[3] = t_cur[0];
        M[4] = Rm[3]; M[5] = Rm[4]; M[6] = Rm[5]; M[7] = t_cur[1];
        M[8] = Rm[6]; M[9] = Rm[7]; M[10] = Rm[8]; M[11] = t_cur[2];
        M[12] = 0; M[13] = 0; M[14] = 0; M[15] = 1;
    }
}

// copy the final (current) keyframe and landmark states to fixed read-back buffers
__global__ void __launch_bounds__(256) lm_export_kernel(const LmCtl* __restrict__ c, Dual q2, Dual t2, Dual pts2, int K, int L, double* __restrict__ qf,
                                                        double* __restrict__ tf, double* __restrict__ pf) {
    const int cur = c->cur & 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 4 * K) qf[i] = q2.p[cur][i];
    if (i < 3 * K) tf[i] = t2.p[cur][i];
    if (i < 3 * L) pf[i] = pts2.p[cur][i];
}

// ---------------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------------
struct Solver {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    unsigned char* d_arena = nullptr;
    size_t arena_cap = 0;
    unsigned char* h_stage = nullptr;  // pinned upload staging
    size_t h_cap = 0;
    double* h_res = nullptr;           // pinned readback
    size_t h_res_cap = 0;
    float last_ms = 0.f;
    int last_launches = 0;
    cudaEvent_t ev_sync = nullptr;
    LmCtl* h_ctl = nullptr;       // pinned mirror of the device control block
    int* h_abort = nullptr;       // pinned, device-visible mirror of the caller's force_stop flag
    int* d_abort = nullptr;
    int chol_cluster = kCholCluster;  // CTAs sharing one factorisation (B200_LBA_CLUSTER overrides: 1, 2, 4 or 8)
    bool host_loop = true;        // false (B200_LBA_GRAPH=1): run the LM loop as one conditional CUDA graph
    // Waiting for the stream (once per LM trial).  Windows are solved many at a time by as many host threads; measured on a 16-core
    // quota with 16 windows in flight next to the front end: spinning in cudaStreamSynchronize gives the best median but starves
    // the thread that feeds the front end every few runs (frames/s halved), a blocking-sync event costs 10-25 %, polling the stream
    // with a 15 us nap in between is within 2 % of spinning and never collapsed.  B200_LBA_WAIT=spin|block|yield|nap overrides.
    int wait_mode = 3;  // 0 spin (cudaStreamSynchronize), 1 blocking event, 2 poll + sched_yield, 3 poll + 15 us sleep
    cudaError_t wait(cudaStream_t st) {
        if (wait_mode == 0) return cudaStreamSynchronize(st);
        if (wait_mode >= 2) {
            cudaError_t e;
            while ((e = cudaStreamQuery(st)) == cudaErrorNotReady) {
                if (wait_mode == 2) sched_yield();
                else std::this_thread::sleep_for(std::chrono::microseconds(15));
            }
            return e;
        }
        cudaError_t e = cudaEventRecord(ev_sync, st);
        return e != cudaSuccess ? e : cudaEventSynchronize(ev_sync);
    }

    int ensure(size_t dev_bytes, size_t host_bytes, size_t res_doubles) {
        if (dev_bytes > arena_cap) {
            if (d_arena) B200_CUDA(cudaFree(d_arena));
            d_arena = nullptr;
            arena_cap = 0;
            B200_CUDA(cudaMalloc(&d_arena, dev_bytes + dev_bytes / 4));
            arena_cap = dev_bytes + dev_bytes / 4;
        }
        if (host_bytes > h_cap) {
            if (h_stage) B200_CUDA(cudaFreeHost(h_stage));
            h_stage = nullptr;
            h_cap = 0;
            B200_CUDA(cudaHostAlloc(&h_stage, host_bytes + host_bytes / 4, cudaHostAllocDefault));
            h_cap = host_bytes + host_bytes / 4;
        }
        if (res_doubles > h_res_cap) {
            if (h_res) B200_CUDA(cudaFreeHost(h_res));
            h_res = nullptr;
            h_res_cap = 0;
            B200_CUDA(cudaHostAlloc(&h_res, sizeof(double) * res_doubles * 2, cudaHostAllocDefault));
            h_res_cap = res_doubles * 2;
        }
        return B200_OK;
    }
};

static std::mutex g_graph_build_mutex;

struct Carver {
    size_t off = 0;
    template <typename T>
    size_t take(size_t n) {
        off = round_up(off, (size_t)256);
        const size_t o = off;
        off += sizeof(T) * std::max<size_t>(n, 1);
        return o;
    }
};

static int solve(Solver& S, const b200_lba_problem_t* P, int iters1, int iters2, volatile uint8_t* force_stop, double* pose_out,
                 double* points_out, uint8_t* outlier_out, b200_lba_stats_t* stats) {
    const int K = P->n_poses, L = P->n_points, E = P->n_edges;
    if (stats) std::memset(stats, 0, sizeof(*stats));
    const bool debug = getenv("B200_LBA_DEBUG") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto ms_since = [](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
    // ---- flatten / sort on the host ------------------------------------------------------------------------------
    std::vector<int> pose_col(K), pt_col(L);
    int Kf = 0, Lf = 0;
    for (int k = 0; k < K; ++k) pose_col[k] = P->pose_fixed[k] ? -1 : Kf++;
    for (int l = 0; l < L; ++l) pt_col[l] = (P->point_fixed && P->point_fixed[l]) ? -1 : Lf++;
    std::vector<int> pt_start(L + 1, 0), order(E);
    for (int e = 0; e < E; ++e) {
        if (P->e_point[e] < 0 || P->e_point[e] >= L || P->e_pose[e] < 0 || P->e_pose[e] >= K || P->e_cam[e] >= P->n_cams) {
            set_error("b200_lba_solve: edge %d references an invalid vertex/camera", e);
            return B200_ERR_INVALID;
        }
        pt_start[P->e_point[e] + 1]++;
    }
    for (int l = 0; l < L; ++l) pt_start[l + 1] += pt_start[l];
    {
        std::vector<int> fill(pt_start.begin(), pt_start.end() - 1);
        for (int e = 0; e < E; ++e) order[fill[P->e_point[e]]++] = e;  // stable: original order within a landmark
    }
    std::unique_ptr<EdgeS[]> edges_buf(new EdgeS[std::max(E, 1)]);  // filled below: no zero-initialisation pass
    EdgeS* edges = edges_buf.get();
    std::vector<unsigned char> robust(E);
    std::vector<int> pose_start(Kf + 1, 0), pose_edges;
    for (int s = 0; s < E; ++s) {
        const int e = order[s];
        EdgeS& d = edges[s];
        d.pose = P->e_pose[e];
        d.pcol = pose_col[d.pose];
        d.point = P->e_point[e];
        d.lcol = pt_col[d.point];
        d.ox = P->e_obs[3 * e]; d.oy = P->e_obs[3 * e + 1]; d.oxr = P->e_obs[3 * e + 2];
        d.inv_sigma_sq = P->e_inv_sigma_sq[e];
        d.delta = P->e_delta[e];
        d.cam = P->e_cam[e];
        d.robust = P->e_robust ? P->e_robust[e] : 1;
        d.can_outlier = P->e_can_be_outlier ? P->e_can_be_outlier[e] : 1;
        d.pad = 0;
        robust[s] = d.robust;
        if (d.pcol >= 0) pose_start[d.pcol + 1]++;
    }
    for (int k = 0; k < Kf; ++k) pose_start[k + 1] += pose_start[k];
    pose_edges.resize(std::max(1, pose_start[Kf]));
    {
        std::vector<int> fill(pose_start.begin(), pose_start.end() - 1);
        for (int s = 0; s < E; ++s)
            if (edges[s].pcol >= 0) pose_edges[fill[edges[s].pcol]++] = s;
    }
    // Schur pair list grouped by upper block (i <= j)
    const int n_blocks = Kf * (Kf + 1) / 2;
    // one pass over the landmarks lists every (a, c) with its block; a counting sort by block then gives the grouped list.  A
    // keyframe observes a landmark once, so a landmark contributes at most one pair to a block and the pairs of a block stay in
    // landmark order whatever the order inside a landmark: its free-keyframe edges are sorted by column first and only the
    // upper-triangular combinations are formed.
    std::vector<int> blk_count(n_blocks + 1, 0);
    size_t pair_bound = 0;
    for (int l = 0; l < L; ++l) {
        const size_t m = (size_t)(pt_start[l + 1] - pt_start[l]);
        pair_bound += m * (m + 1) / 2;
    }
    std::unique_ptr<int2[]> gen_pairs(new int2[pair_bound + 1]);
    std::unique_ptr<int[]> gen_block(new int[pair_bound + 1]);
    size_t n_gen = 0;
    std::vector<int> loc_col, loc_idx;
    loc_col.reserve(64);
    loc_idx.reserve(64);
    for (int l = 0; l < L; ++l) {
        if (pt_col[l] < 0) continue;
        loc_col.clear();
        loc_idx.clear();
        for (int a = pt_start[l]; a < pt_start[l + 1]; ++a) {
            const int pa = edges[a].pcol;
            if (pa < 0) continue;
            size_t pos = loc_col.size();  // insertion sort by column (a handful of entries)
            loc_col.push_back(pa);
            loc_idx.push_back(a);
            while (pos > 0 && loc_col[pos - 1] > pa) {
                loc_col[pos] = loc_col[pos - 1];
                loc_idx[pos] = loc_idx[pos - 1];
                --pos;
            }
            loc_col[pos] = pa;
            loc_idx[pos] = a;
        }
        const int m = (int)loc_col.size();
        for (int x = 0; x < m; ++x) {
            const int pa = loc_col[x];
            const int row = pa * Kf - pa * (pa - 1) / 2 - pa;  // block_id(pa, pc) = row + pc
            for (int y = x; y < m; ++y) {
                const int bid = row + loc_col[y];
                gen_pairs[n_gen] = make_int2(loc_idx[x], loc_idx[y]);
                gen_block[n_gen++] = bid;
                blk_count[bid + 1]++;
            }
        }
    }
    for (int b2 = 0; b2 < n_blocks; ++b2) blk_count[b2 + 1] += blk_count[b2];
    const int n_pairs = blk_count[n_blocks];
    std::vector<int2> pairs(std::max(1, n_pairs));
    {
        std::vector<int> fill(blk_count.begin(), blk_count.end() - 1);
        for (int i = 0; i < n_pairs; ++i) pairs[fill[gen_block[i]]++] = gen_pairs[i];
    }
    std::vector<SchurBlock> blocks(std::max(1, n_blocks));
    std::vector<SchurChunk> chunks;
    std::vector<int> empty_blocks;
    for (int i = 0, b = 0; i < Kf; ++i)
        for (int j = i; j < Kf; ++j, ++b) {
            const int c0 = (int)chunks.size();
            for (int sidx = blk_count[b]; sidx < blk_count[b + 1]; sidx += kSchurChunk)
                chunks.push_back(SchurChunk{sidx, std::min(sidx + kSchurChunk, blk_count[b + 1]), i == j ? 1 : 0, b});
            blocks[b] = SchurBlock{i, j, c0, (int)chunks.size()};
            if (c0 == (int)chunks.size()) empty_blocks.push_back(b);  // no pair falls into this block: it is just Hpp + lambda I or zero
        }
    const int n_chunks = (int)chunks.size();
    // pose-side chunks: kPoseChunk edges of one free keyframe per warp
    std::vector<int2> pose_chunks;          // (start, end) into pose_edges
    std::vector<int> pose_chunk_start(Kf + 1, 0);
    for (int k = 0; k < Kf; ++k) {
        for (int sidx = pose_start[k]; sidx < pose_start[k + 1]; sidx += kPoseChunk)
            pose_chunks.push_back(make_int2(sidx, std::min(sidx + kPoseChunk, pose_start[k + 1])));
        pose_chunk_start[k + 1] = (int)pose_chunks.size();
    }
    const int n_pose_chunks = (int)pose_chunks.size();
    // initial state: util::converter::to_g2o_SE3 (util/converter.cc:17-21)
    std::vector<double> q0(4 * (size_t)std::max(K, 1)), t0(3 * (size_t)std::max(K, 1)), Rt0(12 * (size_t)std::max(K, 1));
    for (int k = 0; k < K; ++k) {
        const double* M = P->pose_cw + 16 * (size_t)k;
        const double R[9] = {M[0], M[1], M[2], M[4], M[5], M[6], M[8], M[9], M[10]};
        rot_to_quat(R, &q0[4 * k]);
        quat_normalize(&q0[4 * k]);
        t0[3 * k] = M[3]; t0[3 * k + 1] = M[7]; t0[3 * k + 2] = M[11];
        quat_to_rot(&q0[4 * k], &Rt0[12 * k]);
        Rt0[12 * k + 9] = M[3]; Rt0[12 * k + 10] = M[7]; Rt0[12 * k + 11] = M[11];
    }
    std::vector<Cam> cams(std::max(1, P->n_cams));
    for (int i = 0; i < P->n_cams; ++i) {
        const b200_camera_t& c = P->cams[i];
        cams[i] = Cam{c.model, c.fx, c.fy, c.cx, c.cy, c.fxb, c.cols, c.rows};
    }

    if (debug) fprintf(stderr, "[lba] host plan %.3f ms (E = %d, pairs = %d, chunks = %d)\n", ms_since(t_begin), E, n_pairs, n_chunks);
    // ---- device arena ----------------------------------------------------------------------------------------------
    const int n = 6 * Kf;
    const int eb = ceil_div(std::max(E, 1), kEdgeThreads), lb = ceil_div(std::max(L, 1), 128), lb2 = ceil_div(std::max(L, 1), 16);
    Carver up;  // uploaded region (mirrors the pinned staging buffer)
    const size_t o_edges = up.take<EdgeS>(E), o_cams = up.take<Cam>(P->n_cams), o_ptstart = up.take<int>(L + 1);
    const size_t o_posestart = up.take<int>(Kf + 1), o_poseedges = up.take<int>(pose_edges.size()), o_robust = up.take<unsigned char>(E);
    const size_t o_posecol = up.take<int>(K), o_blocks = up.take<SchurBlock>(n_blocks), o_pairs = up.take<int2>(n_pairs);
    const size_t o_chunks = up.take<SchurChunk>(n_chunks), o_pchunks = up.take<int2>(n_pose_chunks), o_pcstart = up.take<int>(Kf + 1);
    const size_t o_q0 = up.take<double>(4 * (size_t)K), o_t0 = up.take<double>(3 * (size_t)K), o_Rt0 = up.take<double>(12 * (size_t)K);
    const size_t o_pts0 = up.take<double>(3 * (size_t)L);
    const int n_empty = (int)empty_blocks.size();
    const size_t o_empty = up.take<int>(n_empty), o_blkdone = up.take<int>(n_blocks), o_tickets = up.take<int>(2);  // counters upload as zeros
    const size_t upload_bytes = round_up(up.off, (size_t)256);
    Carver dv;
    dv.off = upload_bytes;
    const size_t o_q1 = dv.take<double>(4 * (size_t)K), o_t1 = dv.take<double>(3 * (size_t)K), o_Rt1 = dv.take<double>(12 * (size_t)K);
    const size_t o_pts1 = dv.take<double>(3 * (size_t)L);
    const size_t o_level = dv.take<unsigned char>(E), o_chi0 = dv.take<double>(E), o_chi1 = dv.take<double>(E);
    const size_t o_Hpl = dv.take<double>(18 * (size_t)E), o_pl = dv.take<double>(9 * (size_t)E);
    const size_t o_Hll = dv.take<double>(6 * (size_t)Lf), o_bl = dv.take<double>(3 * (size_t)Lf), o_Dinv = dv.take<double>(6 * (size_t)Lf);
    const size_t o_part = dv.take<double>(42 * (size_t)n_chunks), o_ppart = dv.take<double>(27 * (size_t)n_pose_chunks), o_Hpp = dv.take<double>(36 * (size_t)Kf), o_bp = dv.take<double>(6 * (size_t)Kf);
    const int ld = n + 2;
    const size_t o_Hs = dv.take<double>((size_t)(n + 1) * ld), o_xp = dv.take<double>(n);
    // readback block: [chi partials eb][diag partials lb][scale partials lb2][result 2]
    const size_t res_n = (size_t)eb + (size_t)lb + (size_t)lb2 + 6;
    const size_t o_res = dv.take<double>(res_n), o_fail = dv.take<int>(1), o_out = dv.take<unsigned char>(E);
    const size_t o_ctl = dv.take<LmCtl>(1), o_qf = dv.take<double>(4 * (size_t)K), o_tf = dv.take<double>(3 * (size_t)K), o_pf = dv.take<double>(3 * (size_t)L);
    int rc = S.ensure(dv.off + 256, upload_bytes, res_n + 36 * (size_t)std::max(Kf, 1));
    if (rc) return rc;
    unsigned char* hs = S.h_stage;
    std::memset(hs, 0, upload_bytes);
    auto put = [&](size_t off, const void* src, size_t bytes) { if (bytes) std::memcpy(hs + off, src, bytes); };
    put(o_edges, edges, sizeof(EdgeS) * E);
    put(o_cams, cams.data(), sizeof(Cam) * P->n_cams);
    put(o_ptstart, pt_start.data(), sizeof(int) * (L + 1));
    put(o_posestart, pose_start.data(), sizeof(int) * (Kf + 1));
    put(o_poseedges, pose_edges.data(), sizeof(int) * pose_edges.size());
    put(o_robust, robust.data(), E);
    put(o_posecol, pose_col.data(), sizeof(int) * K);
    put(o_blocks, blocks.data(), sizeof(SchurBlock) * n_blocks);
    put(o_chunks, chunks.data(), sizeof(SchurChunk) * n_chunks);
    put(o_pchunks, pose_chunks.data(), sizeof(int2) * n_pose_chunks);
    put(o_pcstart, pose_chunk_start.data(), sizeof(int) * (Kf + 1));
    put(o_pairs, pairs.data(), sizeof(int2) * n_pairs);
    put(o_q0, q0.data(), sizeof(double) * 4 * K);
    put(o_t0, t0.data(), sizeof(double) * 3 * K);
    put(o_Rt0, Rt0.data(), sizeof(double) * 12 * K);
    put(o_pts0, P->points, sizeof(double) * 3 * (size_t)L);
    put(o_empty, empty_blocks.data(), sizeof(int) * n_empty);
    if (debug) fprintf(stderr, "[lba] host plan + staging %.3f ms (upload %.2f MB)\n", ms_since(t_begin), upload_bytes / 1e6);
    unsigned char* d = S.d_arena;
    cudaStream_t st = S.stream;
    int launches = 0;
    B200_CUDA(cudaEventRecord(S.ev0, st));
    B200_CUDA(cudaMemcpyAsync(d, hs, upload_bytes, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemsetAsync(d + o_level, 0, E ? E : 1, st));
    B200_CUDA(cudaMemsetAsync(d + o_chi0, 0, sizeof(double) * (E ? E : 1), st));
    B200_CUDA(cudaMemsetAsync(d + o_fail, 0, sizeof(int), st));

    View v{K, L, E, Kf, Lf, (const EdgeS*)(d + o_edges), (const Cam*)(d + o_cams), (const int*)(d + o_ptstart), (const int*)(d + o_posestart),
           (const int*)(d + o_poseedges), d + o_level, d + o_robust};
    double* qs[2] = {(double*)(d + o_q0), (double*)(d + o_q1)};
    double* ts[2] = {(double*)(d + o_t0), (double*)(d + o_t1)};
    double* Rts[2] = {(double*)(d + o_Rt0), (double*)(d + o_Rt1)};
    double* ptss[2] = {(double*)(d + o_pts0), (double*)(d + o_pts1)};
    double* chis[2] = {(double*)(d + o_chi0), (double*)(d + o_chi1)};
    double *Hpl = (double*)(d + o_Hpl), *pl = (double*)(d + o_pl), *Hll = (double*)(d + o_Hll), *bl = (double*)(d + o_bl);
    double *Dinv = (double*)(d + o_Dinv), *part = (double*)(d + o_part), *ppart = (double*)(d + o_ppart), *Hpp = (double*)(d + o_Hpp), *bp = (double*)(d + o_bp);
    double *Hs = (double*)(d + o_Hs), *xp = (double*)(d + o_xp), *res = (double*)(d + o_res);
    const size_t chol_smem = sizeof(double) * ((size_t)kNB * (kNB + 1) + 4 + (size_t)((n + 1 + 3) & ~3) * kNB);
    B200_CUDA(cudaFuncSetAttribute(chol_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chol_smem));
    int* fail = (int*)(d + o_fail);
    double* r_chi = res;                 // eb
    double* r_diag = res + eb;           // lb
    double* r_scale = res + eb + lb;     // lb
    double* r_result = res + eb + lb + lb2;  // 2
    Dual dq{{qs[0], qs[1]}}, dt{{ts[0], ts[1]}}, dRt{{Rts[0], Rts[1]}}, dpts{{ptss[0], ptss[1]}}, dchi{{chis[0], chis[1]}};
    LmCtl* ctl = (LmCtl*)(d + o_ctl);
    LmCtl* hc = S.h_ctl;
    std::memset(hc, 0, sizeof(LmCtl));
    hc->ok = 1;
    *S.h_abort = 0;
    hc->abort_word = force_stop ? S.d_abort : nullptr;
    B200_CUDA(cudaMemcpyAsync(ctl, hc, sizeof(LmCtl), cudaMemcpyHostToDevice, st));
    const bool use_graph = !S.host_loop;
    const int ug = use_graph ? 1 : 0;

    // the kernels of one buildSystem and of one LM trial (identical in graph and host-stepped mode)
    auto tail_of = [&](int ticket_idx, int n_launches, cudaGraphConditionalHandle h_in, cudaGraphConditionalHandle h_out) {
        CtlTail t{};
        t.ctl = ctl;
        t.ticket = (int*)(d + o_tickets) + ticket_idx;
        t.r_chi = r_chi; t.r_diag = r_diag; t.r_scale = r_scale; t.r_result = r_result; t.Hpp = Hpp;
        t.eb = E ? eb : 0; t.lb = lb; t.lb2 = lb2; t.Kf = Kf; t.n_launches = n_launches; t.use_graph = ug;
        t.h_inner = h_in; t.h_outer = h_out;
        return t;
    };
    // computeActiveErrors + buildSystem: 3 launches; the last CTA of the keyframe-side kernel also runs the LM bookkeeping
    auto launch_build = [&](cudaGraphConditionalHandle h_in) -> int {
        if (E) edges_kernel<<<eb, kEdgeThreads, 0, st>>>(v, dRt, dpts, dchi, Hpl, pl, r_chi, 1, 0, nullptr, ctl, CtlTail{});
        points_kernel<<<lb, 128, 0, st>>>(v, pl, Hll, bl, r_diag);
        if (n_pose_chunks) {
            pose_chunks_kernel<<<ceil_div(n_pose_chunks, 4), 128, 0, st>>>(v, (const int2*)(d + o_pchunks), n_pose_chunks, dRt, dpts, ppart, ctl,
                                                                          (const int*)(d + o_pcstart), Hpp, bp, tail_of(0, 3, h_in, h_in));
        } else {
            if (Kf) {
                B200_CUDA(cudaMemsetAsync(Hpp, 0, sizeof(double) * 36 * Kf, st));
                B200_CUDA(cudaMemsetAsync(bp, 0, sizeof(double) * 6 * Kf, st));
            }
            lm_after_build_kernel<<<1, kCtlThreads, 0, st>>>(ctl, r_chi, E ? eb : 0, r_diag, lb, Hpp, Kf, 3, h_in, ug);
        }
        return B200_OK;
    };
    // one LM trial: 5 launches; the last CTA of the trial's chi2 evaluation runs the accept / reject bookkeeping
    auto launch_trial = [&](cudaGraphConditionalHandle h_in, cudaGraphConditionalHandle h_out) -> int {
        if (Lf) dinv_kernel<<<ceil_div(Lf, 128), 128, 0, st>>>(Lf, ctl, Hll, Dinv, fail);
        if (n_blocks)
            schur_chunks_kernel<<<ceil_div(std::max(n_chunks + n_empty, 1), 4), 128, 0, st>>>(
                v, (const SchurChunk*)(d + o_chunks), n_chunks, (const int2*)(d + o_pairs), Hpl, Dinv, bl, part, ctl, (const SchurBlock*)(d + o_blocks),
                (const int*)(d + o_empty), n_empty, (int*)(d + o_blkdone), Hpp, bp, Hs, n, ld);
        {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(S.chol_cluster);
            cfg.blockDim = dim3(kCholThreads);
            cfg.dynamicSmemBytes = chol_smem;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = S.chol_cluster;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            B200_CUDA(cudaLaunchKernelEx(&cfg, chol_solve_kernel, n, ld, Hs, (const LmCtl*)ctl, (const double*)bp, xp, K, (const int*)(d + o_posecol), dq,
                                         dt, dRt, r_result, fail));
        }
        backsub_kernel<<<lb2, 128, 0, st>>>(v, ctl, Dinv, bl, Hpl, xp, dpts, r_scale, fail);
        if (E) {
            edges_kernel<<<eb, kEdgeThreads, 0, st>>>(v, dRt, dpts, dchi, Hpl, pl, r_chi, 0, 1, fail, ctl, tail_of(1, 5, h_in, h_out));
        } else {
            B200_CUDA(cudaMemsetAsync(fail, 0, sizeof(int), st));
            lm_after_trial_kernel<<<1, kCtlThreads, 0, st>>>(ctl, r_chi, 0, r_scale, lb2, r_result, 5, h_in, h_out, ug);
        }
        return B200_OK;
    };
    auto launch_round_tail = [&](int round) -> int {  // chi2 of every active edge at the final state (terminate action's computeActiveErrors)
        if (E) edges_kernel<<<eb, kEdgeThreads, 0, st>>>(v, dRt, dpts, dchi, Hpl, pl, r_chi, 0, 0, nullptr, ctl, CtlTail{});
        lm_round_end_kernel<<<1, kCtlThreads, 0, st>>>(ctl, r_chi, E ? eb : 0, round);
        return B200_OK;
    };
    const int iters[2] = {iters1, iters2};
    int rc2;

    if (use_graph) {
        // ---- one graph: round 1 WHILE { build; WHILE { trial } } -> outliers -> round 2 WHILE { ... } -> report
        // (graph construction is serialised across solver instances: concurrent capture-to-graph + instantiate of conditional
        // graphs from several threads crashed inside the driver on 580.159; execution of the instantiated graphs is concurrent)
        std::unique_lock<std::mutex> build_lock(g_graph_build_mutex);
        cudaGraph_t g = nullptr;
        B200_CUDA(cudaGraphCreate(&g, 0));
        struct GraphGuard {
            cudaGraph_t g;
            std::unique_lock<std::mutex>* lk;
            cudaGraphExec_t ex = nullptr;
            ~GraphGuard() {  // (runs before build_lock's destructor; graph teardown is serialised like construction)
                if (!lk->owns_lock()) lk->lock();
                if (ex) cudaGraphExecDestroy(ex);
                if (g) cudaGraphDestroy(g);
                lk->unlock();
            }
        } guard{g, &build_lock};
        cudaGraphConditionalHandle ho[2], hi[2];
        for (int r = 0; r < 2; ++r) {
            B200_CUDA(cudaGraphConditionalHandleCreate(&ho[r], g, 0, cudaGraphCondAssignDefault));
            B200_CUDA(cudaGraphConditionalHandleCreate(&hi[r], g, 0, cudaGraphCondAssignDefault));
        }
        // a WHILE node appended to the graph that `st` is currently capturing into; returns its (empty) body graph
        auto add_while = [&](cudaGraph_t parent, cudaGraphConditionalHandle hnd, cudaGraph_t* body) -> int {
            cudaStreamCaptureStatus cs;
            const cudaGraphNode_t* deps = nullptr;
            size_t n_deps = 0;
            cudaGraph_t cap = nullptr;
            B200_CUDA(cudaStreamGetCaptureInfo(st, &cs, nullptr, &cap, &deps, &n_deps));
            cudaGraphNodeParams prm = {};
            prm.type = cudaGraphNodeTypeConditional;
            prm.conditional.handle = hnd;
            prm.conditional.type = cudaGraphCondTypeWhile;
            prm.conditional.size = 1;
            cudaGraphNode_t node;
            B200_CUDA(cudaGraphAddNode(&node, parent, deps, n_deps, &prm));
            B200_CUDA(cudaStreamUpdateCaptureDependencies(st, &node, 1, cudaStreamSetCaptureDependencies));
            *body = prm.conditional.phGraph_out[0];
            return B200_OK;
        };
        cudaGraph_t outer_body[2] = {nullptr, nullptr}, inner_body[2] = {nullptr, nullptr}, ended = nullptr;
        B200_CUDA(cudaStreamBeginCaptureToGraph(st, g, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
        for (int r = 0; r < 2; ++r) {
            lm_round_begin_kernel<<<1, 1, 0, st>>>(ctl, iters[r], r, ho[r], 1);
            if (r == 1 && E) outlier_kernel<<<eb, 128, 0, st>>>(v, dRt, dpts, dchi, 0, nullptr, ctl);  // :323-344 (skips itself after an abort)
            if ((rc2 = add_while(g, ho[r], &outer_body[r]))) return rc2;
            if ((rc2 = launch_round_tail(r))) return rc2;
        }
        if (E) outlier_kernel<<<eb, 128, 0, st>>>(v, dRt, dpts, dchi, 1, d + o_out, ctl);  // :354-375
        lm_export_kernel<<<ceil_div(std::max(std::max(4 * K, 3 * L), 1), 256), 256, 0, st>>>(ctl, dq, dt, dpts, K, L, (double*)(d + o_qf), (double*)(d + o_tf),
                                                                                         (double*)(d + o_pf));
        B200_CUDA(cudaStreamEndCapture(st, &ended));
        for (int r = 0; r < 2; ++r) {
            B200_CUDA(cudaStreamBeginCaptureToGraph(st, outer_body[r], nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
            if ((rc2 = launch_build(hi[r]))) return rc2;
            if ((rc2 = add_while(outer_body[r], hi[r], &inner_body[r]))) return rc2;
            B200_CUDA(cudaStreamEndCapture(st, &ended));
            B200_CUDA(cudaStreamBeginCaptureToGraph(st, inner_body[r], nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
            if ((rc2 = launch_trial(hi[r], ho[r]))) return rc2;
            B200_CUDA(cudaStreamEndCapture(st, &ended));
        }
        B200_CUDA(cudaGraphInstantiate(&guard.ex, g, 0));
        B200_CUDA(cudaGraphLaunch(guard.ex, st));
        build_lock.unlock();
        B200_CUDA(cudaMemcpyAsync(hc, ctl, sizeof(LmCtl), cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaEventRecord(S.ev1, st));
        // the caller's flag may be raised by another thread while the graph runs (mapping_module.cc:124): mirror it into the
        // device-visible word that the control kernels test between iterations
        if (force_stop) {
            while (cudaEventQuery(S.ev1) == cudaErrorNotReady) {
                if (*force_stop) *S.h_abort = 1;
                std::this_thread::sleep_for(std::chrono::microseconds(30));
            }
        }
        B200_CUDA(S.wait(st));
    } else {
        // ---- host-stepped: the same kernels, one round trip per loop decision
        auto fetch = [&]() -> int {
            B200_CUDA(cudaMemcpyAsync(hc, ctl, sizeof(LmCtl), cudaMemcpyDeviceToHost, st));
            B200_CUDA(S.wait(st));
            if (force_stop && *force_stop) *S.h_abort = 1;
            return B200_OK;
        };
        for (int r = 0; r < 2; ++r) {
            lm_round_begin_kernel<<<1, 1, 0, st>>>(ctl, iters[r], r, cudaGraphConditionalHandle{}, 0);
            if (r == 1 && E) outlier_kernel<<<eb, 128, 0, st>>>(v, dRt, dpts, dchi, 0, nullptr, ctl);
            if ((rc2 = fetch())) return rc2;
            while (hc->outer_go) {
                if ((rc2 = launch_build(cudaGraphConditionalHandle{}))) return rc2;
                do {
                    if ((rc2 = launch_trial(cudaGraphConditionalHandle{}, cudaGraphConditionalHandle{}))) return rc2;
                    if ((rc2 = fetch())) return rc2;
                } while (hc->inner_go);
            }
            if ((rc2 = launch_round_tail(r))) return rc2;
        }
        if (E) outlier_kernel<<<eb, 128, 0, st>>>(v, dRt, dpts, dchi, 1, d + o_out, ctl);
        lm_export_kernel<<<ceil_div(std::max(std::max(4 * K, 3 * L), 1), 256), 256, 0, st>>>(ctl, dq, dt, dpts, K, L, (double*)(d + o_qf), (double*)(d + o_tf),
                                                                                         (double*)(d + o_pf));
        B200_CUDA(cudaMemcpyAsync(hc, ctl, sizeof(LmCtl), cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaEventRecord(S.ev1, st));
        B200_CUDA(S.wait(st));
    }
    B200_CUDA(cudaGetLastError());
    if (getenv("B200_LBA_DEBUG")) {
        double cyc[6];
        B200_CUDA(cudaMemcpy(cyc, r_result, sizeof(cyc), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[lba] last chol_solve cycles: diag %.0f panel %.0f trailing %.0f backward %.0f (n = %d)\n", cyc[2], cyc[3], cyc[4], cyc[5], n);
    }
    launches = hc->launches + 3;
    if (stats) {
        stats->lambda_init = hc->lambda_init;
        for (int r = 0; r < 2; ++r) {
            stats->iterations[r] = hc->iters_done[r];
            stats->chi2[r] = hc->chi2[r];
            stats->lambda_final[r] = hc->lambda_final[r];
        }
    }
    // terminate_action's gain-threshold stop writes the caller's flag (terminate_action.cc:66-70); an externally raised flag stays up
    if (force_stop && hc->stop_flag) *force_stop = 1;
    // results (exported from whichever buffer ended up current)
    std::vector<unsigned char> out_sorted(std::max(E, 1));
    std::vector<double> qf(4 * (size_t)std::max(K, 1)), tf(3 * (size_t)std::max(K, 1));
    if (E) B200_CUDA(cudaMemcpyAsync(out_sorted.data(), d + o_out, E, cudaMemcpyDeviceToHost, st));
    if (K) {
        B200_CUDA(cudaMemcpyAsync(qf.data(), d + o_qf, sizeof(double) * 4 * K, cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaMemcpyAsync(tf.data(), d + o_tf, sizeof(double) * 3 * K, cudaMemcpyDeviceToHost, st));
    }
    if (L) B200_CUDA(cudaMemcpyAsync(points_out, d + o_pf, sizeof(double) * 3 * (size_t)L, cudaMemcpyDeviceToHost, st));
    B200_CUDA(S.wait(st));
    B200_CUDA(cudaEventElapsedTime(&S.last_ms, S.ev0, S.ev1));
    S.last_launches = launches;
    int n_out = 0;
    for (int s = 0; s < E; ++s) {
        if (outlier_out) outlier_out[order[s]] = out_sorted[s];
        n_out += out_sorted[s];
    }
    if (stats) stats->n_outliers = n_out;
    for (int k = 0; k < K; ++k) {  // util::converter::to_eigen_mat (util/converter.cc:23-25)
        double* M = pose_out + 16 * (size_t)k;
        if (P->pose_fixed[k]) {
            std::memcpy(M, P->pose_cw + 16 * (size_t)k, sizeof(double) * 16);
            continue;
        }
        double R[9];
        quat_to_rot(&qf[4 * k], R);
        M[0] = R[0]; M[1] = R[1]; M[2] = R[2]; M[3] = tf[3 * k];
        M[4] = R[3]; M[5] = R[4]; M[6] = R[5]; M[7] = tf[3 * k + 1];
        M[8] = R[6]; M[9] = R[7]; M[10] = R[8]; M[11] = tf[3 * k + 2];
        M[12] = 0; M[13] = 0; M[14] = 0; M[15] = 1;
    }
    return B200_OK;
}

}  // namespace lba
}  // namespace b200

struct b200_lba_s {
    b200::lba::Solver s;
};

extern "C" {

int b200_lba_create(int device, b200_lba_t* out) {
    if (!out) return B200_ERR_INVALID;
    int rc = b200::require_device(device);
    if (rc) return rc;
    b200_lba_s* h = new (std::nothrow) b200_lba_s();
    if (!h) return B200_ERR_INVALID;
    h->s.device = device;
    // Local BA is the mapping thread's work (mapping_module.cc:63): tracking must not wait for it.  Its ~130 small launches per
    // window are latency-bound (each one leaves most SMs idle), so on the highest stream priority every one of them pre-empts the wide
    // front-end grids for its whole duration (measured: FAST 2.0 -> 2.9 ms per 64 frames with four windows in flight); on the lowest
    // priority its CTAs fill the gaps instead (lowest == the default priority 0 of ordinary streams on this GPU; the range is [0, -5]).
    // B200_LBA_PRIORITY=high|normal|low overrides.
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    int prio = prio_lo;
    if (const char* pe = getenv("B200_LBA_PRIORITY")) prio = (pe[0] == 'l') ? prio_lo : ((pe[0] == 'n') ? 0 : prio_hi);
    cudaError_t e = cudaStreamCreateWithPriority(&h->s.stream, cudaStreamNonBlocking, prio);
    if (e == cudaSuccess) e = cudaEventCreate(&h->s.ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&h->s.ev1);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->s.ev_sync, cudaEventBlockingSync | cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&h->s.h_ctl, sizeof(b200::lba::LmCtl), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&h->s.h_abort, sizeof(int), cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&h->s.d_abort, h->s.h_abort, 0);
    // The conditional-graph driver is opt-in (B200_LBA_GRAPH=1): it is parity-green and slightly faster for one window at a time
    // (4.55 vs 4.94 ms GPU time), but four or more instances executing such graphs concurrently crashed inside driver 580.159
    // (tools/lba_conc.py), and concurrent windows are the normal case here.
    if (const char* cc = getenv("B200_LBA_CLUSTER")) {
        const int c = atoi(cc);
        if (c == 1 || c == 2 || c == 4 || c == 8) h->s.chol_cluster = c;
    }
    if (const char* w = getenv("B200_LBA_WAIT")) h->s.wait_mode = w[0] == 'b' ? 1 : (w[0] == 'y' ? 2 : (w[0] == 'n' ? 3 : 0));  // spin | block | yield | nap
    const char* gm = getenv("B200_LBA_GRAPH");
    h->s.host_loop = !(gm && gm[0] == '1');
    if (e != cudaSuccess) {
        delete h;
        return b200::cuda_fail(e, "stream/event creation", __FILE__, __LINE__);
    }
    *out = h;
    return B200_OK;
}

int b200_lba_destroy(b200_lba_t h) {
    if (!h) return B200_OK;
    cudaSetDevice(h->s.device);
    if (h->s.stream) cudaStreamSynchronize(h->s.stream);
    cudaFree(h->s.d_arena);
    if (h->s.h_stage) cudaFreeHost(h->s.h_stage);
    if (h->s.h_res) cudaFreeHost(h->s.h_res);
    if (h->s.h_ctl) cudaFreeHost(h->s.h_ctl);
    if (h->s.h_abort) cudaFreeHost(h->s.h_abort);
    if (h->s.ev0) cudaEventDestroy(h->s.ev0);
    if (h->s.ev1) cudaEventDestroy(h->s.ev1);
    if (h->s.ev_sync) cudaEventDestroy(h->s.ev_sync);
    if (h->s.stream) cudaStreamDestroy(h->s.stream);
    delete h;
    return B200_OK;
}

int b200_lba_solve(b200_lba_t h, const b200_lba_problem_t* P, int iters1, int iters2, volatile uint8_t* force_stop, double* pose_cw_out,
                   double* points_out, uint8_t* outlier_out, b200_lba_stats_t* stats) {
    if (!h || !P || !pose_cw_out || !points_out) {
        b200::set_error("b200_lba_solve: null argument");
        return B200_ERR_INVALID;
    }
    if (P->n_poses < 0 || P->n_points < 0 || P->n_edges < 0 || P->n_cams < 0 || iters1 < 0 || iters2 < 0
        || (P->n_poses > 0 && (!P->pose_cw || !P->pose_fixed)) || (P->n_points > 0 && !P->points)
        || (P->n_edges > 0 && (!P->e_pose || !P->e_point || !P->e_cam || !P->e_obs || !P->e_inv_sigma_sq || !P->e_delta || !P->cams))) {
        b200::set_error("b200_lba_solve: inconsistent problem description");
        return B200_ERR_INVALID;
    }
    if (6 * (size_t)P->n_poses > 1000) {
        b200::set_error("b200_lba_solve: more than 166 keyframes in one local window is not supported");
        return B200_ERR_INVALID;
    }
    if (force_stop && *force_stop) return B200_ERR_ABORTED;  // local_bundle_adjuster_g2o.cc:308-310
    B200_CUDA(cudaSetDevice(h->s.device));
    return b200::lba::solve(h->s, P, iters1, iters2, force_stop, pose_cw_out, points_out, outlier_out, stats);
}

int b200_pose_optimize(b200_lba_t h, int n_problems, const b200_lba_problem_t* problems, int num_trials_robust, int num_trials, int num_each_iter,
                       double* pose_cw_out, uint8_t* outlier_flags, uint32_t* n_valid) {
    using namespace b200::lba;
    if (!h || n_problems < 0 || num_trials_robust < 0 || num_trials < 0 || num_each_iter < 0) return B200_ERR_INVALID;
    if (n_problems == 0) return B200_OK;
    if (!problems || !pose_cw_out || !n_valid) {
        b200::set_error("b200_pose_optimize: null argument");
        return B200_ERR_INVALID;
    }
    Solver& S = h->s;
    B200_CUDA(cudaSetDevice(S.device));
    size_t total_edges = 0;
    for (int p = 0; p < n_problems; ++p) {
        const b200_lba_problem_t& P = problems[p];
        if (P.n_poses != 1 || P.n_edges < 0 || P.n_points < 0 || P.n_cams < 1 || !P.pose_cw || !P.cams
            || (P.n_edges > 0 && (!P.points || !P.e_point || !P.e_obs || !P.e_inv_sigma_sq || !P.e_delta))) {
            b200::set_error("b200_pose_optimize: problem %d must hold exactly one pose, its observed landmarks and one edge per observation", p);
            return B200_ERR_INVALID;
        }
        for (int e = 0; e < P.n_edges; ++e)
            if (P.e_point[e] < 0 || P.e_point[e] >= P.n_points || (P.e_cam && P.e_cam[e] >= P.n_cams)) {
                b200::set_error("b200_pose_optimize: edge %d of problem %d references an invalid landmark/camera", e, p);
                return B200_ERR_INVALID;
            }
        total_edges += (size_t)P.n_edges;
    }
    if (total_edges > 0 && !outlier_flags) return B200_ERR_INVALID;
    Carver up;
    const size_t o_probs = up.take<PoseProb>(n_problems), o_edges = up.take<PoseEdge>(total_edges);
    const size_t upload_bytes = b200::round_up(up.off, (size_t)256);
    Carver dv;
    dv.off = upload_bytes;
    const size_t o_level = dv.take<unsigned char>(total_edges), o_flags = dv.take<unsigned char>(total_edges);
    const size_t o_pose = dv.take<double>(16 * (size_t)n_problems), o_valid = dv.take<unsigned>(n_problems);
    int rc = S.ensure(dv.off + 256, upload_bytes, 16);
    if (rc) return rc;
    unsigned char* hs = S.h_stage;
    PoseProb* hp = reinterpret_cast<PoseProb*>(hs + o_probs);
    PoseEdge* he = reinterpret_cast<PoseEdge*>(hs + o_edges);
    size_t off = 0;
    for (int p = 0; p < n_problems; ++p) {
        const b200_lba_problem_t& P = problems[p];
        PoseProb pb{};
        pb.n = P.n_edges;
        pb.edge_off = (int)off;
        const b200_camera_t& c = P.cams[(P.e_cam && P.n_edges > 0) ? P.e_cam[0] : 0];
        pb.cam = Cam{c.model, c.fx, c.fy, c.cx, c.cy, c.fxb, c.cols, c.rows};
        const double* M = P.pose_cw;  // util::converter::to_g2o_SE3 (util/converter.cc:17-21)
        const double R[9] = {M[0], M[1], M[2], M[4], M[5], M[6], M[8], M[9], M[10]};
        rot_to_quat(R, pb.q);
        quat_normalize(pb.q);
        pb.t[0] = M[3]; pb.t[1] = M[7]; pb.t[2] = M[11];
        hp[p] = pb;
        for (int e = 0; e < P.n_edges; ++e) {
            PoseEdge pe{};
            const double* pw = P.points + 3 * (size_t)P.e_point[e];
            pe.pw[0] = pw[0]; pe.pw[1] = pw[1]; pe.pw[2] = pw[2];
            pe.ox = P.e_obs[3 * e]; pe.oy = P.e_obs[3 * e + 1]; pe.oxr = P.e_obs[3 * e + 2];
            pe.inv_sigma_sq = P.e_inv_sigma_sq[e];
            pe.delta = P.e_delta[e];
            he[off + e] = pe;
        }
        off += (size_t)P.n_edges;
    }
    unsigned char* d = S.d_arena;
    cudaStream_t st = S.stream;
    B200_CUDA(cudaEventRecord(S.ev0, st));
    B200_CUDA(cudaMemcpyAsync(d, hs, upload_bytes, cudaMemcpyHostToDevice, st));
    pose_optimize_kernel<<<n_problems, kPoseThreads, 0, st>>>((const PoseProb*)(d + o_probs), (const PoseEdge*)(d + o_edges), d + o_level, d + o_flags,
                                                             num_trials_robust, num_trials, num_each_iter, (double*)(d + o_pose),
                                                             (unsigned*)(d + o_valid));
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaEventRecord(S.ev1, st));
    B200_CUDA(cudaMemcpyAsync(pose_cw_out, d + o_pose, sizeof(double) * 16 * (size_t)n_problems, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaMemcpyAsync(n_valid, d + o_valid, sizeof(unsigned) * (size_t)n_problems, cudaMemcpyDeviceToHost, st));
    if (total_edges) B200_CUDA(cudaMemcpyAsync(outlier_flags, d + o_flags, total_edges, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    B200_CUDA(cudaEventElapsedTime(&S.last_ms, S.ev0, S.ev1));
    S.last_launches = 1;
    for (int p = 0; p < n_problems; ++p)  // fewer than five observations: the reference returns before touching the pose (:116-118)
        if (problems[p].n_edges < 5) std::memcpy(pose_cw_out + 16 * (size_t)p, problems[p].pose_cw, sizeof(double) * 16);
    return B200_OK;
}

int b200_lba_last_profile(b200_lba_t h, float* gpu_ms, int* launches) {
    if (!h) return B200_ERR_INVALID;
    if (gpu_ms) *gpu_ms = h->s.last_ms;
    if (launches) *launches = h->s.last_launches;
    return B200_OK;
}

}  // extern "C"
