// abi_common.cu -- error reporting, device checks and pinned-memory helpers of the C ABI.
#include "common.cuh"

namespace b200 {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return B200_ERR_CUDA;
}

int require_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        set_error("no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
        return B200_ERR_CUDA;
    }
    if (device < 0 || device >= n) {
        set_error("device %d out of range (have %d)", device, n);
        return B200_ERR_INVALID;
    }
    int major = 0;
    B200_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10) {
        set_error("device %d has compute capability %d.x; the kernels are built for sm_100a only", device, major);
        return B200_ERR_CUDA;
    }
    B200_CUDA(cudaSetDevice(device));
    return B200_OK;
}

}  // namespace b200

extern "C" {

const char* b200_last_error(void) { return b200::g_err; }

const char* b200_version(void) { return "b200vslam 0.1.0 sm_100a"; }

int b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int b200_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return B200_ERR_INVALID;
    B200_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
    return B200_OK;
}

int b200_host_free(void* ptr) {
    B200_CUDA(cudaFreeHost(ptr));
    return B200_OK;
}

}  // extern "C"
