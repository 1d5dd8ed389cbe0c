// Library plumbing: error text, device queries, frequency-grid check and the
// FMA pipe peak measurement used as the roofline denominator.
#include <atomic>
#include <cmath>

#include "afr_common.cuh"

namespace afr {

static thread_local std::string g_last_error;
static std::atomic<unsigned long long> g_launches{0};
static thread_local int g_fused_path = 0;
static thread_local int g_dft_path = 0;

void retain_pool_memory() {
    static std::atomic<unsigned> done_mask{0};  // one bit per device
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) return;
    const unsigned bit = 1u << dev;
    if (done_mask.load(std::memory_order_relaxed) & bit) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        // keep up to 2 GiB of freed scratch cached between calls (the default threshold, 0,
        // returns everything to the driver at every synchronisation); larger scratch -- y-split
        // partials, wsclean spectra -- goes back, so an allocator sharing the GPU is not starved
        unsigned long long threshold = 2ull << 30;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    }
    done_mask.fetch_or(bit, std::memory_order_relaxed);
}

void note_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

void set_error(const std::string &msg) { g_last_error = msg; }

void note_fused_path(int path) { g_fused_path = path; }
void note_dft_path(int path) { g_dft_path = path; }

int fail(const std::string &msg) {
    g_last_error = msg;
    return 1;
}

namespace {

// Each thread runs kChains independent FMA chains; with 1024 resident threads per
// SM this saturates the FP64 (or FP32) pipe.  2 FLOP per FMA.
template <typename T, int kChains>
__global__ void __launch_bounds__(256) fma_peak_kernel(T *sink, int iters, T a, T b) {
    T acc[kChains];
#pragma unroll
    for (int k = 0; k < kChains; ++k) acc[k] = (T)(threadIdx.x + k);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < kChains; ++k) acc[k] = fma(acc[k], a, b);
    }
    T s = 0;
#pragma unroll
    for (int k = 0; k < kChains; ++k) s += acc[k];
    if (s == (T)123456789) sink[0] = s;  // never true; keeps the chains alive
}

}  // namespace
}  // namespace afr

using namespace afr;

extern "C" int afr_version(void) { return AFR_VERSION; }

extern "C" unsigned long long afr_kernel_launches(void) {
    return g_launches.load(std::memory_order_relaxed);
}

extern "C" const char *afr_last_error(void) { return g_last_error.c_str(); }

extern "C" int afr_last_fused_path(void) { return g_fused_path; }

extern "C" int afr_last_dft_path(void) { return g_dft_path; }

extern "C" int afr_trim_scratch(int device) {
    // hand the stream-ordered scratch this library keeps cached in the device's default CUDA
    // memory pool back to the driver (callers that share the GPU with another allocator)
    cudaMemPool_t pool;
    AFR_CUDA_OK(cudaDeviceGetDefaultMemPool(&pool, device));
    AFR_CUDA_OK(cudaMemPoolTrimTo(pool, 0));
    return 0;
}

extern "C" int afr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int afr_set_device(int device) {
    AFR_CUDA_OK(cudaSetDevice(device));
    return 0;
}

extern "C" int afr_device_info(int device, int *sm_count_, int *cc_major, int *cc_minor,
                               int *clock_khz) {
    int v = 0;
    AFR_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
    if (sm_count_) *sm_count_ = v;
    AFR_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, device));
    if (cc_major) *cc_major = v;
    AFR_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, device));
    if (cc_minor) *cc_minor = v;
    AFR_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, device));
    if (clock_khz) *clock_khz = v;
    return 0;
}

extern "C" int afr_freq_is_uniform(const double *freq, int64_t nchan, double rtol) {
    if (nchan <= 2) return 1;
    const double step = (freq[nchan - 1] - freq[0]) / (double)(nchan - 1);
    double amax = 0.0;
    for (int64_t i = 0; i < nchan; ++i) amax = std::fmax(amax, std::fabs(freq[i]));
    for (int64_t i = 0; i < nchan; ++i) {
        const double want = freq[0] + step * (double)i;
        if (!(std::fabs(freq[i] - want) <= rtol * amax)) return 0;
    }
    return 1;
}

extern "C" int afr_measure_fma_peak(int fp64, int iters, double *flops_per_s, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AFR_REQUIRE(iters > 0 && flops_per_s != nullptr, "afr_measure_fma_peak: bad arguments");
    constexpr int kChains = 8;
    const int blocks = sm_count() * 8;
    void *sink = nullptr;
    AFR_CUDA_OK(cudaMalloc(&sink, 64));
    cudaEvent_t e0, e1;
    AFR_CUDA_OK(cudaEventCreate(&e0));
    AFR_CUDA_OK(cudaEventCreate(&e1));
    float best_ms = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {  // first rep is the warm-up
        AFR_CUDA_OK(cudaEventRecord(e0, stream));
        if (fp64)
            fma_peak_kernel<double, kChains><<<blocks, 256, 0, stream>>>((double *)sink, iters,
                                                                        0.999999, 1e-9);
        else
            fma_peak_kernel<float, kChains><<<blocks, 256, 0, stream>>>((float *)sink, iters,
                                                                       0.999999f, 1e-9f);
        AFR_CUDA_OK(cudaEventRecord(e1, stream));
        AFR_CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0.f;
        AFR_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    AFR_CUDA_OK(cudaGetLastError());
    const double fmas = (double)blocks * 256.0 * (double)kChains * (double)iters;
    *flops_per_s = 2.0 * fmas / ((double)best_ms * 1e-3);
    return 0;
}
