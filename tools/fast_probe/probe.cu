// Probe: the packed arc evaluation (fast_m4) against the candidate test + scalar score (fast_candidates4 / fast_score1) on random
// tiles, on the GPU.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -I../../stella_vslam_b200/csrc probe.cu
#include "../../stella_vslam_b200/csrc/orb_kernels.cu"
#include <cstdio>
#include <cstdlib>
using namespace b200::orb;
__global__ void probe_kernel(const unsigned char* tiles, int n_tiles, int t_low, int* stats, int* first_bad) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const unsigned char* tile = tiles + (size_t)t * 7 * kTilePitch;  // 7 rows; evaluate the word at columns 8..11 of the middle row
    unsigned w[7][3];
    for (int r = 0; r < 7; ++r)
        for (int q = 0; q < 3; ++q) w[r][q] = *reinterpret_cast<const unsigned*>(tile + r * kTilePitch + 4 + 4 * q);
    const unsigned neg_tlow2 = (unsigned)((-t_low) & 0xFFFF) * 0x10001u;
    const unsigned packed = fast_m4(w, neg_tlow2);
    const unsigned cand = fast_candidates4(w, (unsigned)t_low * 0x10001u);
    for (int b = 0; b < 4; ++b) {
        const int want = (packed >> (8 * b)) & 0xFF;
        const int sc = fast_score1(tile + 3 * kTilePitch + 8 + b, t_low);
        atomicAdd(&stats[0], 1);
        if ((cand >> b) & 1) atomicAdd(&stats[1], 1);
        if (want > 0 && !((cand >> b) & 1)) { atomicAdd(&stats[2], 1); atomicCAS(first_bad, -1, t * 4 + b); }
        if (sc != want) { atomicAdd(&stats[3], 1); atomicCAS(first_bad + 1, -1, t * 4 + b); }
    }
}
int main() {
    const int n = 1 << 16;
    unsigned char* h = (unsigned char*)malloc((size_t)n * 7 * kTilePitch);
    srand(1);
    for (int t = 0; t < n; ++t) {
        const int mode = t & 3, base = rand() % 200 + 20;
        for (int i = 0; i < 7 * kTilePitch; ++i)
            h[(size_t)t * 7 * kTilePitch + i] = mode == 0 ? rand() & 255 : (unsigned char)(base + (rand() % (mode == 1 ? 5 : (mode == 2 ? 25 : 60))) - 2);
    }
    unsigned char* d; int *ds, *db;
    cudaMalloc(&d, (size_t)n * 7 * kTilePitch); cudaMalloc(&ds, 16); cudaMalloc(&db, 8);
    cudaMemcpy(d, h, (size_t)n * 7 * kTilePitch, cudaMemcpyHostToDevice);
    for (int t_low : {7, 20}) {
        cudaMemset(ds, 0, 16); cudaMemset(db, 0xFF, 8);
        probe_kernel<<<n / 128, 128>>>(d, n, t_low, ds, db);
        int st[4], bad[2];
        cudaMemcpy(st, ds, 16, cudaMemcpyDeviceToHost); cudaMemcpy(bad, db, 8, cudaMemcpyDeviceToHost);
        printf("t_low %d: pixels %d candidates %d missed %d score mismatches %d (first %d / %d) err=%s\n", t_low, st[0], st[1], st[2], st[3], bad[0], bad[1],
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
