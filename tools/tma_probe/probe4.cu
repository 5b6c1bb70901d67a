#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <cstring>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// MODE 0: descriptor in global memory; MODE 1: grid_constant param, encode via ByVersion(12000)
#if MODE == 0
__global__ void k(const CUtensorMap* tp, int x, int y, unsigned char* out) {
#else
__global__ void k(const __grid_constant__ CUtensorMap t, int x, int y, unsigned char* out) {
    const CUtensorMap* tp = &t;
#endif
    __shared__ __align__(128) unsigned char tile[32 * 64];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(32 * 64) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(tile)),
                     "l"(reinterpret_cast<unsigned long long>(tp)), "r"(x), "r"(y), "r"(smem_u32(&bar)) : "memory");
    }
    __syncthreads();
    asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) out[i] = tile[i];
}
int main() {
    const int w = 320, h = 240, pitch = 320;
    std::vector<unsigned char> img((size_t)pitch * h);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (unsigned char)(i * 7 + (i >> 8));
    unsigned char *d, *o;
    cudaMalloc(&d, img.size()); cudaMalloc(&o, 32 * 64);
    cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t ee = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fp, 12000, cudaEnableDefault, &q);
    printf("entry: %s q=%d fp=%p\n", cudaGetErrorString(ee), (int)q, fp);
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    alignas(64) CUtensorMap tm;
    std::memset(&tm, 0, sizeof(tm));
    cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)h}, strides[1] = {(cuuint64_t)pitch};
    cuuint32_t box[2] = {64, 32}, es[2] = {1, 1};
    CUresult r = ((Fn)fp)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("MODE=%d encode rc %d\n", MODE, (int)r);
    const unsigned long long* wq = reinterpret_cast<const unsigned long long*>(&tm);
    for (int i = 0; i < 16; ++i) printf("  tm[%d]=%016llx\n", i, wq[i]);
#if MODE == 0
    CUtensorMap* dtm; cudaMalloc(&dtm, sizeof(tm)); cudaMemcpy(dtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
    k<<<1, 128>>>(dtm, 16, 8, o);
#else
    k<<<1, 128>>>(tm, 16, 8, o);
#endif
    printf("  run: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
