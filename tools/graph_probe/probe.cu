// Probe: nested WHILE conditional graph nodes populated by stream capture, with a thread-block-cluster kernel and a memset in
// the inner body -- the shape the LM loop of the local BA needs.  Build: nvcc -gencode arch=compute_100a,code=sm_100a probe.cu
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdio>
namespace cg = cooperative_groups;
struct Ctl { int outer, inner, total_inner, cluster_hits; };
__global__ void outer_begin(Ctl* c) { c->inner = 0; }
__global__ void __cluster_dims__(1, 1, 1) dummy() {}
__global__ void cluster_body(Ctl* c) {
    cg::cluster_group cl = cg::this_cluster();
    cl.sync();
    if (threadIdx.x == 0 && cl.block_rank() == 0) atomicAdd(&c->cluster_hits, (int)cl.num_blocks());
}
__global__ void inner_ctl(cudaGraphConditionalHandle h, Ctl* c) {
    c->inner++; c->total_inner++;
    cudaGraphSetConditional(h, c->inner < 1 + (c->outer % 3) ? 1u : 0u);   // 1, 2 or 3 trials
}
__global__ void outer_ctl(cudaGraphConditionalHandle h_outer, cudaGraphConditionalHandle h_inner, Ctl* c, int n) {
    c->outer++;
    cudaGraphSetConditional(h_outer, c->outer < n ? 1u : 0u);
    cudaGraphSetConditional(h_inner, 1u);   // re-arm the inner loop for the next outer iteration
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("FAIL %s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main() {
    Ctl* d; CK(cudaMalloc(&d, sizeof(Ctl))); CK(cudaMemset(d, 0, sizeof(Ctl)));
    int* scratch; CK(cudaMalloc(&scratch, 1024));
    cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaGraph_t g; CK(cudaGraphCreate(&g, 0));
    cudaGraphConditionalHandle h_outer, h_inner;
    CK(cudaGraphConditionalHandleCreate(&h_outer, g, 1, cudaGraphCondAssignDefault));
    CK(cudaGraphConditionalHandleCreate(&h_inner, g, 1, cudaGraphCondAssignDefault));
    cudaGraphNodeParams po = {}; po.type = cudaGraphNodeTypeConditional; po.conditional.handle = h_outer; po.conditional.type = cudaGraphCondTypeWhile; po.conditional.size = 1;
    cudaGraphNode_t n_outer; CK(cudaGraphAddNode(&n_outer, g, nullptr, 0, &po));
    cudaGraph_t g_outer = po.conditional.phGraph_out[0];
    // outer body: outer_begin -> [inner while] -> outer_ctl
    CK(cudaStreamBeginCaptureToGraph(st, g_outer, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    outer_begin<<<1, 1, 0, st>>>(d);
    cudaStreamCaptureStatus cs; const cudaGraphNode_t* deps; size_t ndeps; cudaGraph_t gcap;
    CK(cudaStreamGetCaptureInfo(st, &cs, nullptr, &gcap, &deps, &ndeps));
    cudaGraphNodeParams pi = {}; pi.type = cudaGraphNodeTypeConditional; pi.conditional.handle = h_inner; pi.conditional.type = cudaGraphCondTypeWhile; pi.conditional.size = 1;
    cudaGraphNode_t n_inner; CK(cudaGraphAddNode(&n_inner, g_outer, deps, ndeps, &pi));
    CK(cudaStreamUpdateCaptureDependencies(st, &n_inner, 1, cudaStreamSetCaptureDependencies));
    outer_ctl<<<1, 1, 0, st>>>(h_outer, h_inner, d, 6);
    cudaGraph_t tmp; CK(cudaStreamEndCapture(st, &tmp));
    // inner body: memset + cluster kernel + inner_ctl
    cudaGraph_t g_inner = pi.conditional.phGraph_out[0];
    CK(cudaStreamBeginCaptureToGraph(st, g_inner, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    CK(cudaMemsetAsync(scratch, 0, 1024, st));
    {
        cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(8); cfg.blockDim = dim3(64); cfg.stream = st;
        cudaLaunchAttribute attr[1]; attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = 8; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, cluster_body, d));
    }
    inner_ctl<<<1, 1, 0, st>>>(h_inner, d);
    CK(cudaStreamEndCapture(st, &tmp));
    cudaGraphExec_t ex; CK(cudaGraphInstantiate(&ex, g, 0));
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaMemsetAsync(d, 0, sizeof(Ctl), st));
        CK(cudaGraphLaunch(ex, st)); CK(cudaStreamSynchronize(st));
        Ctl h; CK(cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost));
        printf("rep %d: outer=%d total_inner=%d cluster_hits=%d (expect 6, 1+2+3+1+2+3=12, 96)\n", rep, h.outer, h.total_inner, h.cluster_hits);
    }
    // launch latency of the whole thing
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    CK(cudaMemsetAsync(d, 0, sizeof(Ctl), st));
    cudaEventRecord(e0, st); CK(cudaGraphLaunch(ex, st)); cudaEventRecord(e1, st); CK(cudaStreamSynchronize(st));
    float ms; cudaEventElapsedTime(&ms, e0, e1); printf("graph with 6 outer / 12 inner iterations: %.1f us (%.2f us per kernel-ish node)\n", ms * 1e3, ms * 1e3 / (6 * 2 + 12 * 3));
    return 0;
}
