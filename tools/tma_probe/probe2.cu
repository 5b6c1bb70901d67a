// variants: V=0 single map param; V=1 array param dynamic index; BOXW inner box bytes; RANK 2 or 3
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#ifndef BOXW
#define BOXW 80
#endif
#ifndef BOXH
#define BOXH 72
#endif
#ifndef RANK
#define RANK 3
#endif
struct TmapSet { CUtensorMap m[16]; };
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
#if V == 0
__global__ void k(const __grid_constant__ CUtensorMap t, int level, int x, int y, int z, unsigned char* out) {
    const CUtensorMap* tp = &t;
#else
__global__ void k(const __grid_constant__ TmapSet t, int level, int x, int y, int z, unsigned char* out) {
    const CUtensorMap* tp = &t.m[level];
#endif
    __shared__ __align__(128) unsigned char tile[BOXH * BOXW];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(BOXH * BOXW) : "memory");
#if RANK == 3
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(tile)),
                     "l"(reinterpret_cast<unsigned long long>(tp)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(&bar)) : "memory");
#else
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(tile)),
                     "l"(reinterpret_cast<unsigned long long>(tp)), "r"(x), "r"(y), "r"(smem_u32(&bar)) : "memory");
#endif
    }
    __syncthreads();
    asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < BOXH * BOXW; i += blockDim.x) out[i] = tile[i];
}
int main() {
    const int w = 320, h = 240, pitch = 320, frames = 2;
    std::vector<unsigned char> img((size_t)pitch * h * frames);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (unsigned char)(i * 7 + (i >> 8));
    unsigned char *d, *o;
    cudaMalloc(&d, img.size()); cudaMalloc(&o, BOXH * BOXW);
    cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    TmapSet ts{};
    cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)frames}, strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * h};
    cuuint32_t box[3] = {BOXW, BOXH, 1}, es[3] = {1, 1, 1};
    CUresult r = ((Fn)fp)(&ts.m[3], CU_TENSOR_MAP_DATA_TYPE_UINT8, RANK, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("V=%d RANK=%d BOX=%dx%d encode rc %d\n", V, RANK, BOXW, BOXH, (int)r);
    const int x = 18, y = 19, z = RANK == 3 ? 1 : 0;
#if V == 0
    k<<<1, 256>>>(ts.m[3], 3, x, y, z, o);
#else
    k<<<1, 256>>>(ts, 3, x, y, z, o);
#endif
    cudaError_t e = cudaDeviceSynchronize();
    printf("  run: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<unsigned char> got(BOXH * BOXW);
    cudaMemcpy(got.data(), o, got.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < BOXH; ++r2)
        for (int c = 0; c < BOXW; ++c) {
            const int sx = x + c, sy = y + r2;
            const unsigned char exp = (sx < w && sy < h) ? img[((size_t)z * h + sy) * pitch + sx] : 0;
            bad += got[r2 * BOXW + c] != exp;
        }
    printf("  mismatches %d\n", bad);
    return 0;
}
