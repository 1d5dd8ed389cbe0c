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

// Schedule.  The first two versions had dedicated producer warps (4, then 8) preparing the panels
// while consumer warps multiplied; both stopped at 61 % DMMA-pipe utilisation with the consumers
// waiting on "panel ready" 16-26 % of the time (profiles/ncu_full_fused_dde_mma_r2{a,c}.txt).  Cause:
// the DMMAs keep the FP64 pipe of every SM sub-partition busy (16 cycles each) and the scheduler
// hands the pipe round in turn, so a warp issuing scalar FP64 gets ONE instruction in per rotation
// (~36-50 cycles).  The ~90 FP64 instructions per warp and stage of the phasor + Jones products then
// take longer than the stage's DMMAs -- a latency problem that extra parallelism inside a producer
// warp cannot fix.  So there are no producers: every warp alternates between transforming its share
// of stage s+1 (while its two neighbours on the sub-partition keep the pipe full of DMMAs) and
// multiplying stage s, five shared-memory stages deep so that warps may drift a whole phase apart.
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
    auto b_of = [&](int st) { return smem + st * stage + kKSteps * kstep_bytes; };

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
    const bool shared = p.same_dde && gi0 == gj0 && ni == nj;
    const int per_p = ni * 16, per_q = nj * 16;  // items of one k-step: 8 antennas x 2 sources per group
    const int items_p = kKSteps * per_p, items_q = shared ? 0 : kKSteps * per_q;
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
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(
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
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_addr(b_of(st)) + tid * 16),
                             "l"(reinterpret_cast<const char *>(p.bright) + (s * (long long)p.nchan + f) * 64 +
                                 (tid & 3) * 16));
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_addr(&bars[2 * kNS + st]))
                     : "memory");
    };

    // ---- transform of stage `sg`: Q = k E2 (in place), P = k (E1 B).  Item idx of a panel: k-step
    // idx / per, antenna slot (idx % per) / 2, source parity idx % 2.  The thread <-> item map
    // rotates with the stage so that every warp carries the same share over three stages.
    auto transform = [&](long long sg) {
        const int st = (int)(sg % kNS);
        unsigned char *sbase = smem + st * stage;
        const unsigned char *bb = b_of(st);
        int first = tid + 128 * (int)(sg % 3);
        if (first >= kThreads) first -= kThreads;
        bool waited = false;
        auto do_items = [&](int nitems, int per, int g0, bool is_p) {
            for (int idx = first; idx < nitems; idx += kThreads) {
                const int ks = idx / per, rem = idx - ks * per;
                const int al = rem >> 1, sl = rem & 1, a = g0 * 8 + al;
                const long long s = sg * kSrcPerStage + 2 * ks + sl;
                const bool live = a < nant && s < nsrc;
                const int ac = min(a, nant - 1);
                const long long sc = min(s, nsrc - 1);
                // antenna phasor k_a(s, f) = exp(i psi nu_f), psi = cst (l U_a + m V_a + n W_a): the
                // same operations as the antenna mode of afr_rime_ws.cu -- evaluated BEFORE waiting
                // for the stage's copies
                const double psi = __dmul_rn(p.cst, phase_dot(p.lmn[3 * sc], p.lmn[3 * sc + 1], p.lmn[3 * sc + 2],
                                                              ant_t[3 * ac], ant_t[3 * ac + 1], ant_t[3 * ac + 2], false));
                const C2<double> kk = cis_fast(__dmul_rn(psi, nu));
                const Cd k = {kk.re, kk.im};
                const unsigned row0 = ks * kstep_bytes + (unsigned)(2 * al * kRowBytes);
                const unsigned c0 = chunk_off(2 * sl, al), c1 = chunk_off(2 * sl + 1, al);
                if (!waited) {
                    mbar_wait(&bars[2 * kNS + st], (unsigned)((sg / kNS) & 1));
                    waited = true;
                }
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    // odd antenna slots start with their second row: bank spread of the stores
                    const unsigned row = row0 + (unsigned)((jj ^ (al & 1)) * kRowBytes);
                    Cd x0 = {0.0, 0.0}, x1 = {0.0, 0.0}, m0 = {0.0, 0.0}, m1 = {0.0, 0.0};
                    if (is_p) {
                        if (live) {  // a dead antenna / source was never copied: its rows become zero
                            const unsigned char *src = sbase + (shared ? p_bytes : 0u) + row;
                            x0 = lds_c(src + c0), x1 = lds_c(src + c1);
                            const unsigned char *bm = bb + (2 * ks + sl) * 64;
                            const Cd b0 = lds_c(bm), b1 = lds_c(bm + 16), b2 = lds_c(bm + 32), b3 = lds_c(bm + 48);
                            if (shared) x0 = cmul_(k, x0), x1 = cmul_(k, x1);
                            m0 = cmul2_(x0, b0, x1, b2);
                            m1 = cmul2_(x0, b1, x1, b3);
                            if (!shared) m0 = cmul_(k, m0), m1 = cmul_(k, m1);
                        }
                        if (shared) {
                            sts_c(sbase + p_bytes + row + c0, x0);
                            sts_c(sbase + p_bytes + row + c1, x1);
                        }
                        sts_c(sbase + row + c0, m0);
                        sts_c(sbase + row + c1, m1);
                    } else {
                        unsigned char *q = sbase + p_bytes + row;
                        if (live) x0 = cmul_(k, lds_c(q + c0)), x1 = cmul_(k, lds_c(q + c1));
                        sts_c(q + c0, x0);
                        sts_c(q + c1, x1);
                    }
                }
            }
        };
        do_items(items_p, per_p, gi0, true);
        if (items_q) do_items(items_q, per_q, gj0, false);
        __syncwarp();
        if (p.arrive_all || lane == 0) mbar_arrive(&bars[st]);  // full: this warp's share is in place
    };

    // ---- tiles of this warp: tile `slot * 12 + warp` of the pass; desc = mask | tile_m << 8 |
    // tile_n << 16 (mask: which of the four 8 x 8 blocks = 4 x 4 antennas hold any baseline)
    unsigned desc[kSlots];
#pragma unroll
    for (int sl = 0; sl < kSlots; ++sl) {
        const int ti = sl * kWarps + warp;
        desc[sl] = ti < pass.ntiles ? (unsigned)pass.mask[ti] | ((unsigned)pass.tile_m[ti] << 8) |
                                          ((unsigned)pass.tile_n[ti] << 16)
                                    : 0u;
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

    // one tile, one k-step: 4 LDS.128, 16 DMMA (FULL: all four blocks, no predicates)
    auto tile_step = [&](auto full_tag, int sl, const unsigned char *pb, const unsigned char *qb) {
        constexpr bool FULL = decltype(full_tag)::value;
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
            for (int nn = 0; nn < 2; ++nn)
                if (FULL || (mask_of(sl) & (1u << (2 * mi + nn)))) {
                    dmma(cre[sl][mi][nn], a[mi].x, b[nn].x);
                    dmma(cim[sl][mi][nn], a[mi].y, b[nn].x);
                }
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int nn = 0; nn < 2; ++nn)
                if (FULL || (mask_of(sl) & (1u << (2 * mi + nn)))) {
                    dmma(cre[sl][mi][nn], a[mi].y, b[nn].y);
                    dmma(cim[sl][mi][nn], na[mi], b[nn].y);
                }
    };

    auto consume = [&](long long sg) {
        const int st = (int)(sg % kNS);
        mbar_wait(&bars[st], (unsigned)((sg / kNS) & 1));  // every warp's share of stage sg is in place
#pragma unroll
        for (int ks = 0; ks < kKSteps; ++ks) {
            const unsigned char *pb = smem + st * stage + ks * kstep_bytes + lane_off, *qb = pb + p_bytes;
#pragma unroll
            for (int sl = 0; sl < kSlots; ++sl) {
                if (mask_of(sl) == 0xFu)  // warp-uniform
                    tile_step(std::true_type{}, sl, pb, qb);
                else if (mask_of(sl) != 0u)
                    tile_step(std::false_type{}, sl, pb, qb);
            }
        }
        __syncwarp();
        if (p.arrive_all || lane == 0) mbar_arrive(&bars[kNS + st]);  // empty: done with the buffer
    };

    // ---- pipeline: copies three stages ahead, transform TWO stages ahead, and the three warps of
    // a sub-partition out of step: warps 0-3 multiply first and transform afterwards, warps 4-11
    // transform first.  A warp in its multiply phase can keep the DMMA pipe ~93 % busy on its own
    // (tools/dmma_microbench.cu: 17.3 cycles per DMMA with one warp per sub-partition); what must not
    // happen is all three transforming at once (measured with everyone in the same order: the pipe
    // idles during the common transform phase, 62 % DMMA utilisation).  Two stages of slack mean
    // nobody waits for a neighbour's transform.
    issue(0);
    if (nstage > 1) issue(1);
    if (nstage > 2) issue(2);
    transform(0);
    if (nstage > 1) transform(1);
    for (long long sg = 0; sg < nstage; ++sg) {
        if (sg + 3 < nstage) issue(sg + 3);
        if (warp < 4) {
            consume(sg);
            if (sg + 2 < nstage) transform(sg + 2);
        } else {
            if (sg + 2 < nstage) transform(sg + 2);
            consume(sg);
        }
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
    return kNS * ((size_t)kKSteps * (ni + nj) * 16 * kRowBytes + kSrcPerStage * 64) + 3 * kNS * sizeof(uint64_t);
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
