#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// A: mbarrier only.  B: 1-D bulk copy global->shared with complete_tx.
__global__ void kA(int* out) {
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    if (threadIdx.x == 0) out[0] = 42;
}
__global__ void kB(const unsigned char* src, unsigned char* out) {
    __shared__ __align__(128) unsigned char buf[1024];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1024) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(buf)), "l"(src), "r"(1024),
                     "r"(smem_u32(&bar)) : "memory");
    }
    __syncthreads();
    asm volatile("{\n\t.reg .pred p;\n\tW2: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D2;\n\tbra W2;\n\tD2:\n\t}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = buf[i];
}
int main() {
    int* o; cudaMalloc(&o, 4);
    kA<<<1, 64>>>(o);
    printf("A: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    unsigned char *s, *d; cudaMalloc(&s, 1024); cudaMalloc(&d, 1024); cudaMemset(s, 7, 1024);
    kB<<<1, 64>>>(s, d);
    printf("B: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
