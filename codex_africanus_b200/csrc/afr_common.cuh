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
// every kernel launch of the library is counted (afr_kernel_launches())
void note_launch(int n = 1);
// which kernel the last afr_predict_fused call of this host thread used (AFR_PATH_*)
void note_fused_path(int path);
// which phasor-stream schedule the last DFT-type launch of this host thread used (AFR_DFT_*)
void note_dft_path(int path);

// call right after a <<<...>>> launch: counts it and checks the launch status
#define AFR_LAUNCH_OK()                                                                \
    do {                                                                               \
        ::afr::note_launch();                                                          \
        cudaError_t _e = cudaGetLastError();                                           \
        if (_e != cudaSuccess)                                                         \
            return ::afr::fail(std::string("kernel launch: ") + cudaGetErrorString(_e)); \
    } while (0)

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

// keep up to 2 GiB of freed scratch in the device's default memory pool instead of returning it to
// the driver at every synchronisation (the default release threshold is 0)
void retain_pool_memory();

// stream-ordered scratch allocation
struct Scratch {
    void *ptr = nullptr;
    cudaStream_t stream = nullptr;
    cudaError_t alloc(size_t bytes, cudaStream_t s) {
        stream = s;
        retain_pool_memory();
        return cudaMallocAsync(&ptr, bytes ? bytes : 16, s);
    }
    ~Scratch() {
        if (ptr) cudaFreeAsync(ptr, stream);
    }
};

// --------------------------------------------------------------------------
// shared-memory barriers and TMA bulk copies (warp-specialised kernels)
// --------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_addr(const void *p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n"
        "DONE:\n\t}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;\n" ::"r"(smem_addr(bar)),
                 "r"(bytes)
                 : "memory");
}
// TMA bulk copy global -> shared of `bytes` (multiple of 16, both addresses 16-byte aligned);
// completion is signalled on `bar` as transaction bytes
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
            smem_addr(dst)),
        "l"(src), "r"(bytes), "r"(smem_addr(bar))
        : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_addr(bar)),
                 "r"(bytes)
                 : "memory");
}

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

// exp(i p) in ~25 FP64 instructions, no branches and no calls: Cody-Waite reduction by
// pi/2 in 33-bit pieces (the quotient k is split as k_hi*2^20 + k_lo so that every product
// with the leading piece is exact for |k| < 2^40, i.e. |p| < 1.7e12), then the classic
// minimax kernels on [-pi/4, pi/4] (coefficients of Sun's fdlibm k_sin.c / k_cos.c),
// quadrant fix-up on the integer pipe.  Absolute error 2.2e-16 for |p| <= 1e6 (measured,
// tools/test_cis); beyond that it grows like |p| * 1e-22, far below the |p| * 1e-16
// rounding of the phase itself.  NaN/Inf give NaN.
__device__ __forceinline__ C2<double> cis_fast(double p) {
    const double kMagic = 6755399441055744.0;  // 1.5 * 2^52
    double kd = fma(p, 6.36619772367581382433e-01, kMagic);
    const int k = __double2loint(kd);
    kd -= kMagic;
    // k = kh + kl with kh a multiple of 2^20
    const double kh = (fma(kd, 9.5367431640625e-07, kMagic) - kMagic) * 1048576.0;
    const double kl = kd - kh;
    double r = fma(-kh, 1.57079632673412561417e+00, p);   // pio2_1 (33 bits): exact
    r = fma(-kl, 1.57079632673412561417e+00, r);          // exact
    r = fma(-kd, 6.07710050630396597660e-11, r);          // pio2_2 (33 bits)
    r = fma(-kd, 2.02226624879595063154e-21, r);          // pio2_2t
    const double z = r * r;
    // sin(r)
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01);
    const double sn = fma(z * r, ps, r);
    // cos(r)
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    const double cs = fma(z * z, pc, fma(-0.5, z, 1.0));
    // quadrant: k&1 swaps, sign bits from k
    const double a = (k & 1) ? cs : sn;  // |sin p|
    const double b = (k & 1) ? sn : cs;  // |cos p|
    C2<double> out;
    out.im = __hiloint2double(__double2hiint(a) ^ ((k & 2) << 30), __double2loint(a));
    out.re = __hiloint2double(__double2hiint(b) ^ (((k + 1) & 2) << 30), __double2loint(b));
    return out;
}

}  // namespace afr
