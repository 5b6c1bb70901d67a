// standalone probe of the 3-D u8 TMA tile load used by fast_cells_kernel
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
struct TmapSet { CUtensorMap m[16]; };
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ TmapSet t, int level, int x, int y, int z, unsigned char* out) {
    __shared__ __align__(128) unsigned char tile[72 * 80];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(72 * 80) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(tile)),
                     "l"(reinterpret_cast<unsigned long long>(&t.m[level])), "r"(x), "r"(y), "r"(z), "r"(smem_u32(&bar)) : "memory");
    }
    __syncthreads();
    asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < 72 * 80; i += blockDim.x) out[i] = tile[i];
}
int main() {
    const int w = 320, h = 240, pitch = 320, frames = 2;
    std::vector<unsigned char> img((size_t)pitch * h * frames);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (unsigned char)(i * 7 + (i >> 8));
    unsigned char *d, *o;
    cudaMalloc(&d, img.size()); cudaMalloc(&o, 72 * 80);
    cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    TmapSet ts{};
    cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)frames}, strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * h};
    cuuint32_t box[3] = {80, 72, 1}, es[3] = {1, 1, 1};
    CUresult r = ((Fn)fp)(&ts.m[3], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc %d q %d\n", (int)r, (int)q);
    for (int tc = 0; tc < 3; ++tc) {
        const int x = tc == 0 ? 18 : (tc == 1 ? 274 : 82), y = tc == 0 ? 19 : (tc == 1 ? 200 : 83), z = tc == 2 ? 1 : 0;
        k<<<1, 256>>>(ts, 3, x, y, z, o);
        cudaError_t e = cudaDeviceSynchronize();
        printf("case %d: %s\n", tc, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<unsigned char> got(72 * 80);
        cudaMemcpy(got.data(), o, got.size(), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r2 = 0; r2 < 72; ++r2)
            for (int c = 0; c < 80; ++c) {
                const int sx = x + c, sy = y + r2;
                const unsigned char exp = (sx < w && sy < h && sx >= 0 && sy >= 0) ? img[((size_t)z * h + sy) * pitch + sx] : 0;
                bad += got[r2 * 80 + c] != exp;
            }
        printf("  mismatches %d\n", bad);
    }
    return 0;
}
