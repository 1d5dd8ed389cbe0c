// Full-RIME fused predict with DDEs in antenna-phasor mode as a batched complex FP64 GEMM on the
// FP64 tensor pipe (DMMA, mma.sync.m8n8k4.f64), sm_100a:
//
//   V[t,f][(p,i),(q,j)] = sum_{s,k} P[(p,i),(s,k)] conj(Q[(q,j),(s,k)])
//   P[(p,i),(s,k)] = k_p(s,f) (E1[s,t,p,f] B[s,f])[i,k]      Q[(q,j),(s,k)] = k_q(s,f) E2[s,t,q,f][j,k]
//
// (africanus/rime/predict.py:103-117,193-252 composed with rime/phase.py:20-63, as in
// rime/examples/predict.py:107-134; the per-row phasor K[s,r,f] factorises as k_p conj(k_q) when
// the baseline uvw are differences of per-antenna coordinates, which antenna_uvw_kernel of
// afr_rime_ws.cu checks row by row before this path is allowed.)
//
// Why a GEMM here when the brief says "no dense GEMM is pretended": per (time, channel) the source
// sum IS a rank-2*nsrc update of the (2 nant) x (2 nant) visibility matrix, and the scalar kernel
// (afr_rime_ws.cu, 32 DFMA + 6 LDS.128 per term) is bound by shared-memory wavefronts, not by the
// FP64 pipe (ncu: l1tex 78 %, FP64 47 %).  Measured (tools/dmma_microbench.cu): DMMA.8x8x4 issues
// every 16 cycles per SM sub-partition = 64 FMA lanes/clk/SM = the full 37 TFLOP/s FP64 peak with 4
// operand registers per 8 FMA per thread, where a DFMA with three distinct operands reaches 2/3 of
// it.  A 16 x 16 complex warp tile needs 4 LDS.128 per 16 DMMA: shared memory drops to ~25 %.
//
// CTA = one (time, channel, pass); 12 warps x 3 tiles of 16 x 16 complex outputs (8 x 8 antennas) =
// 36 tiles = the upper triangle of a 64-antenna array in ONE pass (larger arrays: panels of tiles,
// one pass each).  Four sources per stage (= two DMMA k-steps of 4 complex k), five shared-memory
// stages: cp.async of the raw E matrices straight into their final rows three stages ahead, in-place
// scaling by the antenna phasor and P = (k E1) B two stages ahead.  Complex arithmetic on a
// real MMA: a thread loads one complex element (LDS.128) of each operand fragment and issues
//   Re C += Ar Br^T + Ai Bi^T,   Im C += Ai Br^T + (-Ar) Bi^T          (4 DMMA per 8x8x4 complex block).
// The epilogue scatters each thread's (row i of V_pq: two complex values = 32 contiguous bytes) to the
// caller's row order through a (time, antenna1, antenna2) -> row map.
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "afr_dft.cuh"

namespace afr {
namespace {

constexpr int kWarps = 12;      // every warp transforms AND multiplies (see the kernel)
constexpr int kThreads = kWarps * 32;
constexpr int kNS = 5;          // shared-memory stages: consumed | 2 transformed | landing | spare
constexpr int kSlots = 3;       // tiles per warp
constexpr int kRowBytes = 64;   // one panel row: 4 complex k (2 sources x 2) of 16 bytes
constexpr int kKSteps = 2;      // DMMA k-steps (source pairs) per stage
constexpr int kSrcPerStage = 2 * kKSteps;
static_assert(kWarps * kSlots == kDdeMmaMaxTiles, "tiles per pass");

struct Cd {
    double re, im;
};
__device__ __forceinline__ Cd cmul_(Cd a, Cd b) {
    return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
// a * b + c * d with one DMUL + three DFMA per component
__device__ __forceinline__ Cd cmul2_(Cd a, Cd b, Cd c, Cd d) {
    Cd r;
    r.re = fma(-c.im, d.im, fma(c.re, d.re, fma(-a.im, b.im, a.re * b.re)));
    r.im = fma(c.im, d.re, fma(c.re, d.im, fma(a.im, b.re, a.re * b.im)));
    return r;
}
__device__ __forceinline__ Cd lds_c(const unsigned char *p) {
    const double2 v = *reinterpret_cast<const double2 *>(p);
    return {v.x, v.y};
}
__device__ __forceinline__ void sts_c(unsigned char *p, Cd v) {
    *reinterpret_cast<double2 *>(p) = make_double2(v.re, v.im);
}
__device__ __forceinline__ double neg_(double x) {  // sign flip on the integer pipe
    return __hiloint2double(__double2hiint(x) ^ (int)0x80000000, __double2loint(x));
}
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c[0]), "+d"(c[1])
        : "d"(a), "d"(b));
}

// Panel rows are 64 bytes = 4 chunks of one complex value, chunk kc = 2 * (source in k-step) + k.
// The chunk is stored at kc ^ swz(row), swz = bit 2 of the row (= bit 1 of the antenna slot): a
// quarter-warp of a fragment load (rows r, r+1, all four chunks) still covers 128 contiguous
// bytes, and the transform's 16-byte stores of 4 consecutive antenna slots x 2 sources (row parity
// alternating with the slot, see the transform loop) fall into 8 distinct bank groups instead of 2.
__device__ __forceinline__ unsigned chunk_off(int kc, int slot) { return (unsigned)((kc ^ ((slot >> 1) & 1)) * 16); }
__device__ __forceinline__ int st_of(long long sg) { return (int)(sg % kNS); }

// ---- the transform of one "shared panel" item (P = (k E) B, Q = k E) cut into 23 pieces of 2-8
// FP64 instructions, which the multiply phase issues between its groups of four DMMAs.  ptxas does
// not weave two independent instruction streams together by itself (measured: the whole transform
// stayed in front of the 96 DMMAs), and a scalar FP64 instruction only progresses reliably on a
// sub-partition full of DMMAs when it sits in a DMMA warp's own stream.  Same operations, in the same
// order, as cis_fast / phase_dot / the plain transform_item.
struct XformCtx {
    const double *lmn_s, *uvw_a;  // coordinates of the item's source / antenna (clamped indices)
    unsigned char *prow;          // first row of the item in the P panel; the Q panel is +p_bytes
    const unsigned char *bm;      // brightness matrix of the item's source
    double cst, nu;
    unsigned p_bytes, c0, c1, rowa, rowb;  // chunk offsets, byte offsets of the two rows (bank-spread order)
    bool live;
};
struct XformState {
    double l, m, n, u, v, w, t0, t1, ph, kd, r, z, ps, pc, zr, zz;
    int kint;
    Cd k, x0, x1, m0;
};
template <int I>
__device__ __forceinline__ void xform_piece(XformState &x, const XformCtx &c) {
    constexpr double kMagic = 6755399441055744.0;
    auto row_load = [&](unsigned row) {
        const unsigned char *q = c.prow + c.p_bytes + row;
        x.x0 = lds_c(q + c.c0), x.x1 = lds_c(q + c.c1);
        if (!c.live) x.x0 = {0.0, 0.0}, x.x1 = {0.0, 0.0};
        x.x0 = cmul_(x.k, x.x0);
    };
    auto row_q = [&](unsigned row) {
        x.x1 = cmul_(x.k, x.x1);
        unsigned char *q = c.prow + c.p_bytes + row;
        sts_c(q + c.c0, x.x0), sts_c(q + c.c1, x.x1);
    };
    auto row_m0 = [&]() { x.m0 = cmul2_(x.x0, lds_c(c.bm), x.x1, lds_c(c.bm + 32)); };
    auto row_p = [&](unsigned row) {
        const Cd m1 = cmul2_(x.x0, lds_c(c.bm + 16), x.x1, lds_c(c.bm + 48));
        sts_c(c.prow + row + c.c0, x.m0), sts_c(c.prow + row + c.c1, m1);
    };
    if constexpr (I == 0) {
        x.l = c.lmn_s[0], x.m = c.lmn_s[1], x.n = c.lmn_s[2];
        x.u = c.uvw_a[0], x.v = c.uvw_a[1], x.w = c.uvw_a[2];
    } else if constexpr (I == 1) {
        x.t0 = __dmul_rn(x.l, x.u), x.t1 = __dmul_rn(x.m, x.v);
    } else if constexpr (I == 2) {
        x.t0 = __dadd_rn(x.t0, x.t1), x.t1 = __dmul_rn(x.n, x.w);
    } else if constexpr (I == 3) {
        x.ph = __dmul_rn(c.cst, __dadd_rn(x.t0, x.t1));
    } else if constexpr (I == 4) {
        x.ph = __dmul_rn(x.ph, c.nu);
        x.kd = fma(x.ph, 6.36619772367581382433e-01, kMagic);
        x.kint = __double2loint(x.kd);
    } else if constexpr (I == 5) {
        x.kd -= kMagic;
        x.t0 = fma(x.kd, 9.5367431640625e-07, kMagic);
    } else if constexpr (I == 6) {
        x.t0 = (x.t0 - kMagic) * 1048576.0;  // kh
        x.t1 = x.kd - x.t0;                  // kl
    } else if constexpr (I == 7) {
        x.r = fma(-x.t0, 1.57079632673412561417e+00, x.ph);
        x.r = fma(-x.t1, 1.57079632673412561417e+00, x.r);
    } else if constexpr (I == 8) {
        x.r = fma(-x.kd, 6.07710050630396597660e-11, x.r);
        x.r = fma(-x.kd, 2.02226624879595063154e-21, x.r);
    } else if constexpr (I == 9) {
        x.z = x.r * x.r;
        x.ps = fma(x.z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
        x.pc = fma(x.z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    } else if constexpr (I == 10) {
        x.ps = fma(x.z, x.ps, 2.75573137070700676789e-06);
        x.pc = fma(x.z, x.pc, -2.75573143513906633035e-07);
    } else if constexpr (I == 11) {
        x.ps = fma(x.z, x.ps, -1.98412698298579493134e-04);
        x.pc = fma(x.z, x.pc, 2.48015872894767294178e-05);
    } else if constexpr (I == 12) {
        x.ps = fma(x.z, x.ps, 8.33333333332248946124e-03);
        x.pc = fma(x.z, x.pc, -1.38888888888741095749e-03);
    } else if constexpr (I == 13) {
        x.ps = fma(x.z, x.ps, -1.66666666666666324348e-01);
        x.pc = fma(x.z, x.pc, 4.16666666666666019037e-02);
        x.zr = x.z * x.r, x.zz = x.z * x.z;
    } else if constexpr (I == 14) {
        const double sn = fma(x.zr, x.ps, x.r);
        const double cs = fma(x.zz, x.pc, fma(-0.5, x.z, 1.0));
        const int k = x.kint;
        const double a = (k & 1) ? cs : sn, b = (k & 1) ? sn : cs;
        x.k.im = __hiloint2double(__double2hiint(a) ^ ((k & 2) << 30), __double2loint(a));
        x.k.re = __hiloint2double(__double2hiint(b) ^ (((k + 1) & 2) << 30), __double2loint(b));
        if (!c.live) x.k = {0.0, 0.0};  // a dead antenna / source: its rows become zero
    } else if constexpr (I == 15) {
        row_load(c.rowa);
    } else if constexpr (I == 16) {
        row_q(c.rowa);
    } else if constexpr (I == 17) {
        row_m0();
    } else if constexpr (I == 18) {
        row_p(c.rowa);
    } else if constexpr (I == 19) {
        row_load(c.rowb);
    } else if constexpr (I == 20) {
        row_q(c.rowb);
    } else if constexpr (I == 21) {
        row_m0();
    } else if constexpr (I == 22) {
        row_p(c.rowb);
    }
}
constexpr int kXformPieces = 23;

// Schedule.  Measured facts behind it (tools/dmma_mix_microbench.cu, profiles/dmma_mix_microbench_r2.txt):
//   * two or more warps issuing DMMAs on an SM sub-partition keep its FP64 pipe 99.7 % busy -- and
//     STARVE any other warp's scalar FP64 instructions (58-700 DFMAs got through in 5e5 cycles);
//   * one DMMA warp alone reaches 81-84 % (19-20 cycles per DMMA);
//   * scalar FP64 instructions INSIDE a DMMA warp's own instruction stream cost 3-4 pipe cycles each.
// The first versions had dedicated producer warps (4, then 8) preparing the panels, then symmetric
// warps alternating whole transform / multiply phases: all stopped at 60-62 % DMMA utilisation
// because the ~90 scalar FP64 instructions per item (phasor range reduction and polynomials, k E,
// (k E) B) only progressed while the DMMA warps were idle.  Hence: no producers, and every warp
// carries its share of the transform of stage s+2 IN ITS OWN STREAM, interleaved with the DMMAs of
// stage s (in-order issue guarantees the scalar work progresses; the DMMAs hide its latency).
__global__ void __launch_bounds__(kThreads, 1) fused_dde_mma_kernel(const DdeMmaParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f = blockIdx.x, t = blockIdx.y;
    const DdeMmaPass &pass = p.passes[blockIdx.z];
    const int gi0 = pass.gi0, ni = pass.ni, gj0 = pass.gj0, nj = pass.nj;
    const int nant = (int)p.nant;
    // stage = kKSteps x { P panel | Q panel } | B of the stage's sources
    const unsigned p_bytes = (unsigned)ni * 16 * kRowBytes, q_bytes = (unsigned)nj * 16 * kRowBytes;
    const unsigned kstep_bytes = p_bytes + q_bytes;
    const unsigned stage = kKSteps * kstep_bytes + kSrcPerStage * 64;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kNS * stage);  // full, empty, landed [kNS]
    // a 2x2 zero matrix: the "brightness" of a source past the end of the list (the last stage of an
    // nsrc that is not a multiple of four).  Its B slot is never copied, and 0 * (whatever shared
    // memory held) is NaN often enough -- found by tools/fuzz_fused.py with 7 and 19 sources
    unsigned char *zero64 = smem + kNS * stage + 16 * ((3 * kNS * sizeof(uint64_t) + 15) / 16);
    auto b_of = [&](int st) { return smem + st * stage + kKSteps * kstep_bytes; };

    if (tid < 4) reinterpret_cast<double2 *>(zero64)[tid] = make_double2(0.0, 0.0);
    if (tid == 0) {
        for (int i = 0; i < kNS; ++i) {
            // full / empty: one elected lane per warp (after __syncwarp); AFR_SANITIZE=1: every lane
            mbar_init(&bars[i], p.arrive_all ? kThreads : kWarps);
            mbar_init(&bars[kNS + i], p.arrive_all ? kThreads : kWarps);
            mbar_init(&bars[2 * kNS + i], kThreads);  // landed: the cp.async of every thread
        }
    }
    __syncthreads();

    const long long nsrc = p.nsrc;
    const long long nstage = (nsrc + kSrcPerStage - 1) / kSrcPerStage;
    // one E matrix serves both operands when E1 is E2 and the panel is on the diagonal
    const bool shared_rt = p.same_dde && gi0 == gj0 && ni == nj;
    const bool shared = shared_rt;
    const int per_p = ni * 16, per_q = nj * 16;  // items of one k-step: 8 antennas x 2 sources per group
    const int items_p = kKSteps * per_p, items_q = shared ? 0 : kKSteps * per_q;
    const int items = items_p + items_q;
    const double nu = p.freq[f];
    const double *ant_t = p.ant_uvw + (long long)t * nant * 3;
    const long long mat_stride_s = (long long)p.ntime * nant * p.nchan * 64;  // bytes between sources
    const char *e1_tf = reinterpret_cast<const char *>(p.dde1) + ((long long)t * nant * p.nchan + f) * 64;
    const char *e2_tf = reinterpret_cast<const char *>(p.dde2) + ((long long)t * nant * p.nchan + f) * 64;
    const long long ant_stride = (long long)p.nchan * 64;

    // ---- raw E matrices of stage `sg` -> their final panel rows (16-byte cp.async), B beside them
    auto issue = [&](long long sg) {
        const int st = (int)(sg % kNS);
        if (sg >= kNS) mbar_wait(&bars[kNS + st], (unsigned)(((sg - kNS) / kNS) & 1));  // buffer released
        auto copy_items = [&](int nitems, int per, int g0, const char *src_tf, unsigned panel_off) {
            const unsigned dst0 = smem_addr(smem + st * stage) + panel_off;
            for (int idx = tid; idx < nitems; idx += kThreads) {
                const int ks = idx / per, rem = idx - ks * per;
                const int al = rem >> 1, sl = rem & 1, a = g0 * 8 + al;
                const long long s = sg * kSrcPerStage + 2 * ks + sl;
                if (a < nant && s < nsrc) {
                    const char *src = src_tf + s * mat_stride_s + a * ant_stride;
                    const unsigned dst = dst0 + ks * kstep_bytes + (unsigned)(2 * al * kRowBytes);
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16;\n" ::"r"(
                                             dst + j * kRowBytes + chunk_off(2 * sl + k, al)),
                                         "l"(src + (2 * j + k) * 16));
                }
            }
        };
        if (shared) {
            copy_items(items_p, per_p, gj0, e2_tf, p_bytes);
        } else {
            copy_items(items_p, per_p, gi0, e1_tf, 0u);
            copy_items(items_q, per_q, gj0, e2_tf, p_bytes);
        }
        if (tid < kSrcPerStage * 4) {
            const long long s = sg * kSrcPerStage + (tid >> 2);
            if (s < nsrc)
                asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16;\n" ::"r"(smem_addr(b_of(st)) + tid * 16),
                             "l"(reinterpret_cast<const char *>(p.bright) + (s * (long long)p.nchan + f) * 64 +
                                 (tid & 3) * 16));
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_addr(&bars[2 * kNS + st]))
                     : "memory");
    };

    // ---- one item of the transform of stage `sg`: Q = k E2 (in place) or P = k (E1 B).  Item idx:
    // [0, items_p) the P panel, then the Q panel; inside a panel: k-step idx / per, antenna slot
    // (idx % per) / 2, source parity idx % 2.  Branch-free (dead antennas / sources are masked), so
    // that the compiler can weave it into the DMMA stream of the multiply phase.
    auto transform_item = [&](auto fast_tag, long long sg, int idx, bool store) {
        // FAST: the panel is "shared" (every item is a P item whose k E also is the Q panel) -- known at
        // compile time, so the item is one basic block
        constexpr bool FAST = decltype(fast_tag)::value;
        const int st = (int)(sg % kNS);
        unsigned char *sbase = smem + st * stage;
        const bool is_p = FAST || idx < items_p;
        const bool shared = FAST || shared_rt;
        const int li = is_p ? idx : idx - items_p, per = is_p ? per_p : per_q, g0 = is_p ? gi0 : gj0;
        const int ks = li / per, rem = li - ks * per;
        const int al = rem >> 1, sl = rem & 1, a = g0 * 8 + al;
        const long long s = sg * kSrcPerStage + 2 * ks + sl;
        const bool live = a < nant && s < nsrc;
        const int ac = min(a, nant - 1);
        const long long sc = min(s, nsrc - 1);
        // antenna phasor k_a(s, f) = exp(i psi nu_f), psi = cst (l U_a + m V_a + n W_a): the same
        // operations as the antenna mode of afr_rime_ws.cu
        const double psi = __dmul_rn(p.cst, phase_dot(p.lmn[3 * sc], p.lmn[3 * sc + 1], p.lmn[3 * sc + 2],
                                                      ant_t[3 * ac], ant_t[3 * ac + 1], ant_t[3 * ac + 2], false));
        const C2<double> kk = cis_fast(__dmul_rn(psi, nu));
        const Cd k = {live ? kk.re : 0.0, live ? kk.im : 0.0};  // a dead item's rows become zero
        const unsigned row0 = ks * kstep_bytes + (unsigned)(2 * al * kRowBytes);
        const unsigned c0 = chunk_off(2 * sl, al), c1 = chunk_off(2 * sl + 1, al);
        const unsigned char *bm = live ? b_of(st) + (2 * ks + sl) * 64 : zero64;
        // in-place panel of the raw matrix: Q (shared or a Q item) or P
        unsigned char *raw = sbase + ((shared || !is_p) ? p_bytes : 0u);
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
            // odd antenna slots start with their second row: bank spread of the stores
            const unsigned row = row0 + (unsigned)((jj ^ (al & 1)) * kRowBytes);
            Cd x0 = lds_c(raw + row + c0), x1 = lds_c(raw + row + c1);
            if (!live) x0 = {0.0, 0.0}, x1 = {0.0, 0.0};  // never copied: whatever the buffer holds
            if (is_p) {
                const Cd b0 = lds_c(bm), b1 = lds_c(bm + 16), b2 = lds_c(bm + 32), b3 = lds_c(bm + 48);
                Cd m0, m1;
                if (shared) {
                    x0 = cmul_(k, x0), x1 = cmul_(k, x1);
                    if (store) sts_c(raw + row + c0, x0), sts_c(raw + row + c1, x1);
                    m0 = cmul2_(x0, b0, x1, b2);
                    m1 = cmul2_(x0, b1, x1, b3);
                } else {
                    m0 = cmul_(k, cmul2_(x0, b0, x1, b2));
                    m1 = cmul_(k, cmul2_(x0, b1, x1, b3));
                }
                if (store) sts_c(sbase + row + c0, m0), sts_c(sbase + row + c1, m1);
            } else {
                x0 = cmul_(k, x0), x1 = cmul_(k, x1);
                if (store) sts_c(raw + row + c0, x0), sts_c(raw + row + c1, x1);
            }
        }
    };
    // The thread <-> item map rotates with the stage, so that with 256 items (64 antennas) every
    // warp carries an item in two stages out of three and every sub-partition two items per stage.
    auto first_item = [&](long long sg) {
        int first = tid + 128 * (int)(sg % 3);
        return first >= kThreads ? first - kThreads : first;
    };
    auto wait_landed = [&](long long sg) {
        mbar_wait(&bars[2 * kNS + (int)(sg % kNS)], (unsigned)((sg / kNS) & 1));
    };
    auto publish = [&](long long sg) {  // full: this warp's share of stage sg is in place
        __syncwarp();
        if (p.arrive_all || lane == 0) mbar_arrive(&bars[(int)(sg % kNS)]);
    };
    auto transform_all = [&](long long sg) {  // prologue only
        wait_landed(sg);
        for (int idx = first_item(sg); idx < items; idx += kThreads) transform_item(std::false_type{}, sg, idx, true);
        publish(sg);
    };

    // ---- tiles of this warp: tile `slot * 12 + warp` of the pass, (m0, n0) = first complex row in
    // the P / Q panel.  A tile is multiplied whole (all four 8 x 8 blocks) so that the multiply phase
    // is one basic block; `mask` (which 4 x 4-antenna blocks hold baselines) only gates the epilogue.
    unsigned desc[kSlots];
    int my_tiles = 0;
#pragma unroll
    for (int sl = 0; sl < kSlots; ++sl) {
        const int ti = sl * kWarps + warp;
        desc[sl] = 0u;
        if (ti < pass.ntiles) {
            desc[sl] = (unsigned)pass.mask[ti] | ((unsigned)pass.tile_m[ti] << 8) | ((unsigned)pass.tile_n[ti] << 16);
            my_tiles = sl + 1;
        }
    }
    auto mask_of = [&](int sl) { return desc[sl] & 0xFu; };
    auto moff_of = [&](int sl) { return ((desc[sl] >> 8) & 0xFFu) * (16u * kRowBytes); };
    auto noff_of = [&](int sl) { return (desc[sl] >> 16) * (16u * kRowBytes); };
    // fragment element of this lane: row lane / 4 of an 8-row block, chunk lane % 4 (swizzled by
    // bit 2 of the row = bit 4 of the lane)
    const unsigned lane_off = (unsigned)((lane >> 2) * kRowBytes + (((lane & 3) ^ ((lane >> 4) & 1)) * 16));

    double cre[kSlots][2][2][2], cim[kSlots][2][2][2];
#pragma unroll
    for (int sl = 0; sl < kSlots; ++sl)
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int nn = 0; nn < 2; ++nn)
                cre[sl][mi][nn][0] = cre[sl][mi][nn][1] = cim[sl][mi][nn][0] = cim[sl][mi][nn][1] = 0.0;

    // one tile, one k-step: 4 LDS.128, 16 DMMA
    auto tile_step = [&](int sl, const unsigned char *pb, const unsigned char *qb) {
        double2 a[2], b[2];
        const unsigned char *pa = pb + moff_of(sl), *qa = qb + noff_of(sl);
        a[0] = *reinterpret_cast<const double2 *>(pa);
        a[1] = *reinterpret_cast<const double2 *>(pa + 8 * kRowBytes);
        b[0] = *reinterpret_cast<const double2 *>(qa);
        b[1] = *reinterpret_cast<const double2 *>(qa + 8 * kRowBytes);
        const double na[2] = {neg_(a[0].x), neg_(a[1].x)};
        // independent accumulators first, their second products afterwards
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int nn = 0; nn < 2; ++nn) {
                dmma(cre[sl][mi][nn], a[mi].x, b[nn].x);
                dmma(cim[sl][mi][nn], a[mi].y, b[nn].x);
            }
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int nn = 0; nn < 2; ++nn) {
                dmma(cre[sl][mi][nn], a[mi].y, b[nn].y);
                dmma(cim[sl][mi][nn], na[mi], b[nn].y);
            }
    };

    // Multiply stage sg (NT tiles of this warp) with the transform of this thread's first item of
    // stage sg + 2 placed in the middle of the same basic block (HAS: the warp carries items).
    auto body = [&](auto nt_tag, auto has_tag, long long sg) {
        constexpr int NT = decltype(nt_tag)::value;
        constexpr int HAS = decltype(has_tag)::value;  // 0: no item, 1: generic item, 2: shared-panel item
        const int st = (int)(sg % kNS);
        const unsigned char *pb0 = smem + st * stage + lane_off;
        const int idx = first_item(sg + 2);
#pragma unroll
        for (int ks = 0; ks < kKSteps; ++ks) {
            const unsigned char *pb = pb0 + ks * kstep_bytes, *qb = pb + p_bytes;
#pragma unroll
            for (int sl = 0; sl < NT; ++sl) {
                if (HAS == 2 && ks == 0 && sl == 0) transform_item(std::true_type{}, sg + 2, idx < items ? idx : 0, idx < items);
                tile_step(sl, pb, qb);
                if (HAS == 1 && ks == 0 && sl == 0) transform_item(std::false_type{}, sg + 2, idx < items ? idx : 0, idx < items);
            }
        }
    };

    // FAST multiply phase: the warp's NT tiles of stage sg as groups of four DMMAs with the pieces
    // of this thread's shared-panel item of stage sg + 2 in between (one basic block).
    auto body_fast = [&](auto nt_tag, auto abl_tag, long long sg) {
        constexpr int NT = decltype(nt_tag)::value;
        constexpr int ABL = decltype(abl_tag)::value;  // diagnostics: 1 = no transform, 2 = no DMMA
        constexpr int G = NT * kKSteps * 4;  // groups
        const unsigned char *pb0 = smem + st_of(sg) * stage + lane_off;
        const long long sg2 = sg + 2;
        const int idx = first_item(sg2);  // < items: the panel's item count is a multiple of 32
        const int ks2 = idx / per_p, rem = idx - ks2 * per_p;
        const int al = rem >> 1, sl2 = rem & 1, a = gi0 * 8 + al;
        const long long s = sg2 * kSrcPerStage + 2 * ks2 + sl2;
        XformCtx xc;
        xc.live = a < nant && s < nsrc;
        xc.lmn_s = p.lmn + 3 * min(s, nsrc - 1);
        xc.uvw_a = ant_t + 3 * min(a, nant - 1);
        xc.prow = smem + st_of(sg2) * stage + ks2 * kstep_bytes + (unsigned)(2 * al * kRowBytes);
        xc.bm = xc.live ? b_of(st_of(sg2)) + (2 * ks2 + sl2) * 64 : zero64;
        xc.cst = p.cst, xc.nu = nu, xc.p_bytes = p_bytes;
        xc.c0 = chunk_off(2 * sl2, al), xc.c1 = chunk_off(2 * sl2 + 1, al);
        xc.rowa = (unsigned)((al & 1) * kRowBytes), xc.rowb = (unsigned)(((al & 1) ^ 1) * kRowBytes);
        XformState xs;
        double2 fa[2], fb[2];
        double na[2];
        auto pieces = [&](auto self, auto lo_tag, auto hi_tag) {
            constexpr int LO = decltype(lo_tag)::value, HI = decltype(hi_tag)::value;
            if constexpr (LO < HI) {
                xform_piece<LO>(xs, xc);
                self(self, std::integral_constant<int, LO + 1>{}, hi_tag);
            }
        };
        auto group = [&](auto self, auto g_tag) {
            constexpr int g = decltype(g_tag)::value;
            constexpr int ks = g / (NT * 4), sl = (g / 4) % NT, q = g % 4, mi = q & 1;
            if constexpr (q == 0) {
                const unsigned char *pa = pb0 + ks * kstep_bytes + moff_of(sl);
                const unsigned char *qa = pb0 + ks * kstep_bytes + p_bytes + noff_of(sl);
                fa[0] = *reinterpret_cast<const double2 *>(pa);
                fa[1] = *reinterpret_cast<const double2 *>(pa + 8 * kRowBytes);
                fb[0] = *reinterpret_cast<const double2 *>(qa);
                fb[1] = *reinterpret_cast<const double2 *>(qa + 8 * kRowBytes);
                na[0] = neg_(fa[0].x), na[1] = neg_(fa[1].x);
            }
            if constexpr (ABL >= 2) {
                cre[sl][mi][0][0] += fa[mi].x * fb[0].x + na[mi];
            } else if constexpr (q < 2) {
                dmma(cre[sl][mi][0], fa[mi].x, fb[0].x);
                dmma(cim[sl][mi][0], fa[mi].y, fb[0].x);
                dmma(cre[sl][mi][1], fa[mi].x, fb[1].x);
                dmma(cim[sl][mi][1], fa[mi].y, fb[1].x);
            } else {
                dmma(cre[sl][mi][0], fa[mi].y, fb[0].y);
                dmma(cim[sl][mi][0], na[mi], fb[0].y);
                dmma(cre[sl][mi][1], fa[mi].y, fb[1].y);
                dmma(cim[sl][mi][1], na[mi], fb[1].y);
            }
            if constexpr (ABL != 1 && ABL != 3)
                pieces(pieces, std::integral_constant<int, g * kXformPieces / G>{},
                       std::integral_constant<int, (g + 1) * kXformPieces / G>{});
            if constexpr (g + 1 < G) self(self, std::integral_constant<int, g + 1>{});
        };
        group(group, std::integral_constant<int, 0>{});
    };

    // ---- pipeline: copies three stages ahead, transform two stages ahead
    for (int i = 0; i < 3; ++i)
        if (i < nstage) issue(i);
    transform_all(0);
    if (nstage > 1) transform_all(1);
    for (long long sg = 0; sg < nstage; ++sg) {
        if (sg + 3 < nstage) issue(sg + 3);
        const bool ahead = sg + 2 < nstage;
        // warp-uniform: does this warp carry items of stage sg + 2?  (idx < items for its lane 0 ..
        // lane 31 range: the map is contiguous in tid, so test the warp's first lane)
        const int w0 = first_item(sg + 2) - lane;
        const bool has = ahead && w0 < items;
        mbar_wait(&bars[st_of(sg)], (unsigned)((sg / kNS) & 1));  // every warp's share of stage sg is in place
        if (has) wait_landed(sg + 2);
#define AFR_BODY(NT)                                                                       \
    if (has && shared && p.ablate == 1)                                                    \
        body_fast(std::integral_constant<int, NT>{}, std::integral_constant<int, 1>{}, sg); \
    else if (has && shared && p.ablate == 2)                                               \
        body_fast(std::integral_constant<int, NT>{}, std::integral_constant<int, 2>{}, sg); \
    else if (has && shared && p.ablate == 3)                                               \
        body_fast(std::integral_constant<int, NT>{}, std::integral_constant<int, 3>{}, sg); \
    else if (has && shared)                                                                \
        body_fast(std::integral_constant<int, NT>{}, std::integral_constant<int, 0>{}, sg); \
    else if (has)                                                                          \
        body(std::integral_constant<int, NT>{}, std::integral_constant<int, 1>{}, sg);     \
    else                                                                                   \
        body(std::integral_constant<int, NT>{}, std::integral_constant<int, 0>{}, sg)
        if (my_tiles == 3) {
            AFR_BODY(3);
        } else if (my_tiles == 2) {
            AFR_BODY(2);
        } else if (my_tiles == 1) {
            AFR_BODY(1);
        }
#undef AFR_BODY
        if (has) {
            // further items of a large panel, and every item of a warp without tiles (small arrays)
            for (int idx = first_item(sg + 2) + (my_tiles ? kThreads : 0); idx < items; idx += kThreads)
                transform_item(std::false_type{}, sg + 2, idx, true);
        }
        __syncwarp();
        if (p.arrive_all || lane == 0) mbar_arrive(&bars[kNS + st_of(sg)]);  // empty: done with the buffer
        if (ahead) publish(sg + 2);
    }

    // ---- epilogue: this lane holds V_pq[i][0..1] of every block: p = row / 2, i = row % 2,
    // q = first column / 2 -- 32 contiguous bytes of the caller's row (time, p, q)
    const int32_t *map_t = p.rowmap + (long long)t * nant * nant;
#pragma unroll
    for (int sl = 0; sl < kSlots; ++sl) {
        const int ti = sl * kWarps + warp;
        if (ti >= pass.ntiles) continue;
        const int row0 = gi0 * 16 + pass.tile_m[ti] * 16, col0 = gj0 * 16 + pass.tile_n[ti] * 16;
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int nn = 0; nn < 2; ++nn) {
                if (!(mask_of(sl) & (1u << (2 * mi + nn)))) continue;
                const int row = row0 + 8 * mi + (lane >> 2), col = col0 + 8 * nn + 2 * (lane & 3);
                const int pa = row >> 1, i = row & 1, qa = col >> 1;
                if (pa < nant && qa < nant) {
                    const int r = map_t[(long long)pa * nant + qa];
                    if (r >= 0) {
                        double2 *o = reinterpret_cast<double2 *>(p.out + ((long long)r * p.nchan + f) * 8 + i * 4);
                        o[0] = make_double2(cre[sl][mi][nn][0], cim[sl][mi][nn][0]);
                        o[1] = make_double2(cre[sl][mi][nn][1], cim[sl][mi][nn][1]);
                    }
                }
            }
    }
}

// rowmap[t][a1][a2] = row (the caller's order), dup[0] set when two rows share (t, a1, a2);
// used4[(a1 / 4) * n4 + a2 / 4] = 1 where any timestep has a baseline of that 4 x 4 antenna block
__global__ void baseline_map_kernel(const int32_t *time_index, const int32_t *ant1, const int32_t *ant2,
                                    long long nrow, long long ntime, long long nant, int n4, int32_t *rowmap,
                                    uint8_t *used4, int *dup) {
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (r >= nrow) return;
    const long long t = time_index[r], a1 = ant1[r], a2 = ant2[r];
    if (t < 0 || t >= ntime || a1 < 0 || a1 >= nant || a2 < 0 || a2 >= nant) {
        atomicOr(dup, 2);
        return;
    }
    const int old = atomicCAS(&rowmap[(t * nant + a1) * nant + a2], -1, (int)r);
    if (old != -1) atomicOr(dup, 1);
    used4[(a1 >> 2) * n4 + (a2 >> 2)] = 1;
}

}  // namespace

size_t dde_mma_smem_bytes(int ni, int nj) {
    return kNS * ((size_t)kKSteps * (ni + nj) * 16 * kRowBytes + kSrcPerStage * 64) +
           16 * ((3 * kNS * sizeof(uint64_t) + 15) / 16) + 64;  // stages | mbarriers | zero matrix
}

int launch_baseline_map(const int32_t *time_index, const int32_t *ant1, const int32_t *ant2, int64_t nrow,
                        int64_t ntime, int64_t nant, int32_t *rowmap, uint8_t *used4, int *dup,
                        cudaStream_t stream) {
    const int n4 = (int)((nant + 3) / 4);
    AFR_CUDA_OK(cudaMemsetAsync(rowmap, 0xFF, sizeof(int32_t) * (size_t)(ntime * nant * nant), stream));
    AFR_CUDA_OK(cudaMemsetAsync(used4, 0, (size_t)n4 * n4, stream));
    AFR_CUDA_OK(cudaMemsetAsync(dup, 0, sizeof(int), stream));
    if (nrow <= 0) return 0;
    baseline_map_kernel<<<(unsigned)((nrow + 255) / 256), 256, 0, stream>>>(time_index, ant1, ant2, nrow, ntime,
                                                                            nant, n4, rowmap, used4, dup);
    AFR_LAUNCH_OK();
    return 0;
}

// Passes from the used 4 x 4-antenna blocks (host copy): panels of `ps` tile rows x `ps` tile
// columns (a tile = 8 x 8 antennas); every panel pair with a used tile becomes one pass per 36 of
// its used tiles.  64 antennas, a1 < a2: one pass of 36 tiles.
std::vector<DdeMmaPass> dde_mma_passes(const std::vector<uint8_t> &used4, int64_t nant) {
    const int n4 = (int)((nant + 3) / 4), n8 = (int)((nant + 7) / 8);
    const int ps = n8 <= 8 ? 8 : 6;
    std::vector<DdeMmaPass> passes;
    auto u4 = [&](int i, int j) { return i < n4 && j < n4 && used4[(size_t)i * n4 + j] != 0; };
    for (int pi = 0; pi < n8; pi += ps)
        for (int pj = 0; pj < n8; pj += ps) {
            DdeMmaPass cur{};
            auto start = [&]() {
                cur = DdeMmaPass{};
                cur.gi0 = pi, cur.ni = std::min(ps, n8 - pi);
                cur.gj0 = pj, cur.nj = std::min(ps, n8 - pj);
            };
            start();
            for (int mt = 0; mt < std::min(ps, n8 - pi); ++mt)
                for (int nt = 0; nt < std::min(ps, n8 - pj); ++nt) {
                    unsigned m = 0;
                    for (int mi = 0; mi < 2; ++mi)
                        for (int nn = 0; nn < 2; ++nn)
                            if (u4(2 * (pi + mt) + mi, 2 * (pj + nt) + nn)) m |= 1u << (2 * mi + nn);
                    if (!m) continue;
                    if (cur.ntiles == kDdeMmaMaxTiles) {
                        passes.push_back(cur);
                        start();
                    }
                    cur.tile_m[cur.ntiles] = (uint8_t)mt;
                    cur.tile_n[cur.ntiles] = (uint8_t)nt;
                    cur.mask[cur.ntiles] = (uint8_t)m;
                    ++cur.ntiles;
                }
            if (cur.ntiles) passes.push_back(cur);
        }
    return passes;
}

int launch_fused_dde_mma(DdeMmaParams p, const std::vector<DdeMmaPass> &passes, cudaStream_t stream) {
    if (passes.empty() || p.nsrc <= 0) return 0;
    Scratch dpass;
    AFR_CUDA_OK(dpass.alloc(sizeof(DdeMmaPass) * passes.size(), stream));
    AFR_CUDA_OK(cudaMemcpyAsync(dpass.ptr, passes.data(), sizeof(DdeMmaPass) * passes.size(),
                                cudaMemcpyHostToDevice, stream));
    p.passes = (const DdeMmaPass *)dpass.ptr;
    p.arrive_all = (getenv("AFR_SANITIZE") && atoi(getenv("AFR_SANITIZE")) != 0) ? 1 : 0;
    p.ablate = getenv("AFR_MMA_ABLATE") ? atoi(getenv("AFR_MMA_ABLATE")) : 0;  // diagnostics only: wrong results
    int nmax = 0;
    for (const auto &q : passes) nmax = std::max(nmax, q.ni + q.nj);
    const size_t smem = dde_mma_smem_bytes(nmax, 0);
    AFR_REQUIRE(p.ntime <= 65535 && passes.size() <= 65535, "afr_predict_fused: grid too large");
    AFR_CUDA_OK(cudaFuncSetAttribute(fused_dde_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)p.nchan, (unsigned)p.ntime, (unsigned)passes.size());
    fused_dde_mma_kernel<<<grid, kThreads, smem, stream>>>(p);
    AFR_LAUNCH_OK();
    return 0;
}

}  // namespace afr
