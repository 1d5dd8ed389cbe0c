// Shared device/host helpers for libafricanus_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/africanus_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libafricanus_b200 targets sm_100a (B200) only"
#endif

namespace afr {

// africanus/constants/consts.py:6-9
constexpr double kLightSpeed = 2.99792458e8;
constexpr double kTwoPiOverC = 2.0 * 3.141592653589793 / kLightSpeed;

// --------------------------------------------------------------------------
// error plumbing (thread-local text behind afr_last_error())
// --------------------------------------------------------------------------
void set_error(const std::string &msg);
int fail(const std::string &msg);

#define AFR_CUDA_OK(expr)                                                              \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess)                                                         \
            return ::afr::fail(std::string(#expr) + ": " + cudaGetErrorString(_e));    \
    } while (0)

#define AFR_REQUIRE(cond, msg)                 \
    do {                                       \
        if (!(cond)) return ::afr::fail(msg);  \
    } while (0)

inline int sm_count() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess)
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}

// stream-ordered scratch allocation
struct Scratch {
    void *ptr = nullptr;
    cudaStream_t stream = nullptr;
    cudaError_t alloc(size_t bytes, cudaStream_t s) {
        stream = s;
        return cudaMallocAsync(&ptr, bytes ? bytes : 16, s);
    }
    ~Scratch() {
        if (ptr) cudaFreeAsync(ptr, stream);
    }
};

// --------------------------------------------------------------------------
// device math
// --------------------------------------------------------------------------
template <typename T>
struct C2 {
    T re, im;
};

// The phase argument follows the reference's evaluation order exactly:
//   real_phase = constant * (l*u + m*v + n*w)   (phase.py:49, kernels.py:57,128)
// with explicitly rounded multiplies/adds (no FMA contraction).  When both lm
// and uvw were float32 the reference evaluates l*u + m*v in float32.
__device__ __forceinline__ double phase_dot(double x0, double x1, double x2, double y0,
                                            double y1, double y2, bool f32dot) {
    double a;
    if (f32dot) {
        float fa = __fadd_rn(__fmul_rn((float)x0, (float)y0), __fmul_rn((float)x1, (float)y1));
        a = (double)fa;
    } else {
        a = __dadd_rn(__dmul_rn(x0, y0), __dmul_rn(x1, y1));
    }
    return __dadd_rn(a, __dmul_rn(x2, y2));
}

template <typename T>
__device__ __forceinline__ C2<T> cmul(C2<T> a, C2<T> b) {
    C2<T> r;
    r.re = a.re * b.re - a.im * b.im;
    r.im = a.re * b.im + a.im * b.re;
    return r;
}

__device__ __forceinline__ C2<double> cis(double p) {
    C2<double> r;
    sincos(p, &r.im, &r.re);
    return r;
}

}  // namespace afr
