// Library-wide plumbing: error text, device properties, the device-side fault flag.
#include <stdarg.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace cvae {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_add_fetch(&g_launches, 1ULL, __ATOMIC_RELAXED); }
unsigned long long launches() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

bool pdl_enabled() {
    static const bool on = !(getenv("CVAE_PDL") && atoi(getenv("CVAE_PDL")) == 0);
    return on;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// Dynamic shared-memory opt-in.  cudaFuncAttributeMaxDynamicSharedMemorySize is a property of (function, device), not
// of the calling thread: the cache is process-wide, keyed by both, only ever raised, and guarded by a mutex, so
// PyTorch's autograd thread and the main thread (or two batch sizes) can never lower each other's setting.
int opt_in_smem(const void* func, size_t bytes) {
    static std::mutex mu;
    static std::unordered_map<unsigned long long, size_t> configured;
    int dev = 0;
    CVAE_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    size_t& have = configured[((unsigned long long)(uintptr_t)func << 6) ^ (unsigned long long)(dev & 63)];
    if (bytes > have) {
        CVAE_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        have = bytes;
    }
    return CVAE_OK;
}

__device__ int g_fault_flag = 0;

int* fault_flag() {
    static thread_local int* ptr[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!ptr[dev]) {
        void* p = nullptr;
        if (cudaGetSymbolAddress(&p, g_fault_flag) != cudaSuccess) return nullptr;
        ptr[dev] = (int*)p;
    }
    return ptr[dev];
}

}  // namespace cvae

extern "C" const char* cvae_last_error(void) { return cvae::g_err; }
extern "C" int cvae_version(void) { return 100; }
extern "C" int64_t cvae_launch_count(void) { return (int64_t)cvae::launches(); }

extern "C" int cvae_check_device_fault(void* stream) {
    int* p = cvae::fault_flag();
    CVAE_REQUIRE(p != nullptr, CVAE_ECUDA, "fault flag unavailable");
    int v = 0;
    CVAE_CUDA(cudaMemcpyAsync(&v, p, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CVAE_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (v) {
        int zero = 0;
        CVAE_CUDA(cudaMemcpyAsync(p, &zero, sizeof(int), cudaMemcpyHostToDevice, (cudaStream_t)stream));
        CVAE_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
        cvae::set_error("device-side pipeline fault: a bounded mbarrier wait expired");
        return CVAE_EDEVICE;
    }
    return CVAE_OK;
}
