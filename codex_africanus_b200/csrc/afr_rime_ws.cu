// Full-RIME fused predict with DDEs, warp-specialised (sm_100a):
//
//   V[r,f] += sum_s K[s,r,f] * E1[s,t,a1,f] * B[s,f] * E2[s,t,a2,f]^H        (2x2, complex128)
//
// (africanus/rime/predict.py:103-117,193-252 composed with rime/phase.py:20-63; the
// composition is the one of rime/examples/predict.py:107-134.)
//
// A CTA owns one timestep and 512 rows x 4 channels (or 2048 rows x 1 channel in antenna
// mode); lane <-> row, 16 consumer warps; the 64 accumulator doubles of a thread stay in
// registers for the whole source loop.  Per source the 4 PRODUCER warps
//   * fetch E[s,t,:,f0:f0+4] for ALL antennas (and B[s,f0:f0+4]) with 16-byte cp.async tracked
//     by an mbarrier, into a padded shared-memory layout (antenna stride 272 B, so lanes
//     reading consecutive antennas hit distinct banks and lanes reading the same antenna
//     broadcast);
//   * precombine A_p = E1_p * B_s (valid because K is a scalar: E1 (K B) E2^H = K (E1 B) E2^H),
//     in place when E1 != E2;
//   * and supply the phasors, in one of two forms:
//       ROW mode  - each consumer thread computes its row's anchor exp(i phi nu_f0) and channel
//                   step exp(i phi dnu) itself (phi from the row's uvw exactly as the reference
//                   rounds it), multiplies the 2x2 product by the phasor and advances it with
//                   the three-term recurrence.  (The first version had the 4 producer warps
//                   compute the 1024 sincos per source: they could not keep up, 96 Gterms/s.)
//       ANT mode  - when the baseline uvw of a timestep are differences of per-antenna
//                   coordinates (checked on the device, see antenna_uvw_kernel), K factorises as
//                   k_p conj(k_q) and is folded into the antenna matrices: A_p <- k_p A_p,
//                   E_q <- k_q E_q.  64 phasors per source instead of 512, and the consumer is
//                   left with ONE 2x2 product and the accumulate per term.
// Consumers wait on the stage's "full" mbarrier, do register-only FP64 work fed by LDS.128,
// and release the stage with one arrive per warp.  Three stages: one being consumed, one
// being prepared, one in flight from L2/HBM.
#include <algorithm>
#include <cstdlib>

#include <cub/device/device_radix_sort.cuh>

#include "afr_dft.cuh"

namespace afr {
namespace {

constexpr int kConsWarps = 16;
constexpr int kProdWarps = 4;
constexpr int kConsThreads = kConsWarps * 32;
constexpr int kNS = 3;         // pipeline stages
constexpr int kMatBytes = 64;  // one 2x2 complex128 matrix
constexpr int kNTP = kProdWarps * 32;

struct Cd {
    double re, im;
};
__device__ __forceinline__ Cd cmul_(Cd a, Cd b) {
    return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
__device__ __forceinline__ Cd cmulc_(Cd a, Cd b) {  // a * conj(b)
    return {a.re * b.re + a.im * b.im, a.im * b.re - a.re * b.im};
}
__device__ __forceinline__ Cd cadd_(Cd a, Cd b) { return {a.re + b.re, a.im + b.im}; }
// acc += a * conj(b) / acc += a * b as four chained DFMA (no separate DMUL / DADD: the consumer
// loop is FP64-issue and register-file bound, and a DFMA costs no more than either of those)
__device__ __forceinline__ void cfmac_(Cd &acc, Cd a, Cd b) {
    acc.re = fma(a.re, b.re, acc.re);
    acc.im = fma(a.im, b.re, acc.im);
    acc.re = fma(a.im, b.im, acc.re);
    acc.im = fma(-a.re, b.im, acc.im);
}
__device__ __forceinline__ void cfma_(Cd &acc, Cd a, Cd b) {
    acc.re = fma(a.re, b.re, acc.re);
    acc.im = fma(a.re, b.im, acc.im);
    acc.re = fma(-a.im, b.im, acc.re);
    acc.im = fma(a.im, b.re, acc.im);
}

__device__ __forceinline__ Cd lds_c(const unsigned char *p) {
    const double2 v = *reinterpret_cast<const double2 *>(p);
    return {v.x, v.y};
}
__device__ __forceinline__ void sts_c(unsigned char *p, Cd v) {
    *reinterpret_cast<double2 *>(p) = make_double2(v.re, v.im);
}

// shared-memory stride of one antenna: FT matrices + 16 bytes, so that lanes reading
// consecutive antennas with LDS.128 fall into distinct bank groups (8 antennas cover the 32
// banks) and lanes reading the same antenna broadcast
__host__ __device__ constexpr int ant_stride(int ft) { return ft * kMatBytes + 16; }

constexpr int kPlaneBytes = 192;   // two frequency planes x (re, im, |.|) x 4 correlations of one (antenna, channel)
constexpr int kPlaneStride = 224;  // their pitch in shared memory: 128 k + 32, so that the 16-byte loads of a quarter-warp
                                   // (4 items x 2 rows) fall into distinct bank groups (192: 8-way conflicts, ncu 6.8e8)
size_t stage_bytes(int na, int ft, bool ant, bool sample = false) {
    // E2 (or E) | E1 -> A (or A) | B | (sampling: the raw planes)
    (void)ant;
    return 2 * (size_t)na * ant_stride(ft) + (size_t)ft * kMatBytes + (sample ? (size_t)na * ft * kPlaneStride : 0);
}

// FT channels and RPT rows per consumer thread (FT * RPT = 4: 64 accumulator doubles).
//   <4,1>: a CTA owns 512 rows x 4 channels -- the shape the per-row phasor recurrence needs;
//   <1,4>: a CTA owns 2048 rows x 1 channel -- antenna mode: a MeerKAT timestep (2016
//          baselines) is ONE row tile, so every E element is fetched, scaled and precombined
//          exactly once per (source, time, channel) instead of once per 512-row tile.
template <bool EXACT, bool ANT, int FT, int RPT, bool SAMPLE = false>
__global__ void __launch_bounds__((kConsWarps + kProdWarps) * 32, 1)
    fused_dde_ws_kernel(const DdeWsParams p) {
    static_assert(ANT || RPT == 1, "per-row phasors advance along the thread's channel run");
    static_assert(!SAMPLE || (ANT && !EXACT), "in-kernel beam sampling runs in antenna mode");
    static_assert(FT == 1 || FT == 2 || FT == 4, "FT");
    constexpr int AS = ant_stride(FT);
    constexpr int kRows = kConsThreads * RPT;
    constexpr int LFT = FT == 4 ? 2 : (FT == 2 ? 1 : 0);
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int na = (int)p.nant;
    const int t = blockIdx.y;
    const int f0 = blockIdx.z * FT;
    const long long rbeg = p.row_start[t] + (long long)blockIdx.x * kRows;
    const long long rend = min((long long)p.row_start[t + 1], rbeg + kRows);
    if (rbeg >= rend) return;

    const size_t mat_region = (size_t)na * AS;
    const size_t stage = 2 * mat_region + FT * kMatBytes + (SAMPLE ? (size_t)na * FT * kPlaneStride : 0);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kNS * stage);  // full, empty, landed [kNS]
    double *fq = reinterpret_cast<double *>(bars + 3 * kNS);            // [FT]
    auto pl_of = [&](int st) { return smem + st * stage + 2 * mat_region + FT * kMatBytes; };  // SAMPLE: raw planes
    auto e2_of = [&](int st) { return smem + st * stage; };
    auto a_of = [&](int st) { return smem + st * stage + mat_region; };
    auto b_of = [&](int st) { return smem + st * stage + 2 * mat_region; };

    if (tid == 0) {
        for (int i = 0; i < kNS; ++i) {
            mbar_init(&bars[i], kNTP);             // full: every producer thread
            // empty: one elected lane per consumer warp (after __syncwarp); with AFR_SANITIZE=1
            // every lane arrives, which compute-sanitizer's racecheck can follow
            mbar_init(&bars[kNS + i], p.arrive_all ? kConsWarps * 32 : kConsWarps);
            mbar_init(&bars[2 * kNS + i], kNTP);   // landed: the cp.async of every producer thread
        }
    }
    double *fw = fq + FT;  // [FT] (SAMPLE) weight of the lower frequency plane
    if (tid < FT) fq[tid] = p.freq[min(f0 + tid, p.nchan - 1)];
    int *fg = reinterpret_cast<int *>(fw + FT);  // [FT] (SAMPLE) index of the lower frequency plane
    if (SAMPLE && tid < FT) {
        fw[tid] = p.fd[3 * min(f0 + tid, p.nchan - 1) + 1];
        fg[tid] = (int)p.fd[3 * min(f0 + tid, p.nchan - 1) + 2];
    }
    __syncthreads();

    const int valid_ch = min(FT, p.nchan - f0);
    const long long nsrc = p.nsrc;
    double dnu = 0.0;
    if (p.nchan > 1) dnu = (p.freq[p.nchan - 1] - p.freq[0]) / (double)(p.nchan - 1);
    const double nu0 = p.freq[min(f0, p.nchan - 1)];

    if (warp >= kConsWarps) {
        // =============================== PRODUCERS ===============================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;\n");
        const int ptid = tid - kConsWarps * 32;
        // E[s,t,:,f0:f0+FT] (both sides) and B[s,f0:f0+FT] -> stage s % kNS by 16-byte cp.async
        // spread over the 128 producer threads; completion is tracked by the stage's "landed"
        // mbarrier (cp.async.mbarrier.arrive.noinc, one arrival per producer thread).  TMA bulk
        // copies were measured first: the tile is 65-129 pieces of 64-256 B per source, and one
        // bulk copy per piece costs more issue time than the whole consume phase.
        const int gran_valid = valid_ch * 4;  // 16-byte granules per antenna that exist
        const long long astride = (long long)p.nchan * kMatBytes;  // bytes between antennas
        auto issue = [&](long long s) {
            const int st = (int)(s % kNS);
            const long long base = (((s * p.ntime + t) * p.nant) * (long long)p.nchan + f0) * kMatBytes;
            const char *src2 = reinterpret_cast<const char *>(p.dde2) + base;
            const char *src1 = reinterpret_cast<const char *>(p.dde1) + base;
            const unsigned d2 = smem_addr(e2_of(st)), d1 = smem_addr(a_of(st));
            if constexpr (SAMPLE) {
                // the two planes (192 contiguous bytes) of every (antenna, channel) of this tile
                const char *psrc = reinterpret_cast<const char *>(p.planes) +
                                   ((s * p.ntime + t) * p.nant) * (long long)p.nud * 96;
                const unsigned dp = smem_addr(pl_of(st));
                for (int g = ptid; g < na * FT * 12; g += kNTP) {
                    const int item = g / 12, c = g - item * 12;
                    const int a = item >> LFT, fl = item & (FT - 1);
                    if (fl < valid_ch) {
                        const int gl = fg[fl];  // (a global load here sat in front of every copy: ncu long-scoreboard)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dp + item * kPlaneStride + c * 16),
                                     "l"(psrc + ((long long)a * p.nud + gl) * 96 + c * 16));
                    }
                }
            }
            for (int g = ptid; !SAMPLE && g < na * FT * 4; g += kNTP) {
                const int a = g >> (LFT + 2), c = g & (FT * 4 - 1);
                if (c < gran_valid) {
                    const unsigned doff = (unsigned)(a * AS + c * 16);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d2 + doff),
                                 "l"(src2 + a * astride + c * 16));
                    if (!p.same_dde)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d1 + doff),
                                     "l"(src1 + a * astride + c * 16));
                }
            }
            if (ptid < gran_valid)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(
                                 smem_addr(b_of(st)) + ptid * 16),
                             "l"(reinterpret_cast<const char *>(p.bright) +
                                 (s * (long long)p.nchan + f0) * kMatBytes + ptid * 16));
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(
                             smem_addr(&bars[2 * kNS + st]))
                         : "memory");
        };
        for (long long s = 0; s < kNS - 1 && s < nsrc; ++s) issue(s);

        const double *ant_t = ANT ? p.ant_uvw + (long long)t * p.nant * 3 : nullptr;

        for (long long s = 0; s < nsrc; ++s) {
            const int st = (int)(s % kNS);
            const unsigned par = (unsigned)((s / kNS) & 1);
            // the stage must have been released by the consumers of source s - kNS
            if (s >= kNS) mbar_wait(&bars[kNS + st], (unsigned)(((s - kNS) / kNS) & 1));
            const double sl = p.lmn[3 * s], sm = p.lmn[3 * s + 1], sn = p.lmn[3 * s + 2];
            mbar_wait(&bars[2 * kNS + st], par);  // E (and B) of source s have landed
            // ---- half-items (antenna a, channel fl, output row h): row h of A_a = E1_a * B_s,
            // and in antenna mode row h of k_a A_a and of k_a E2_a
            const unsigned char *bsm = b_of(st);
            unsigned char *e2 = e2_of(st), *am = a_of(st);
            for (int idx = ptid; idx < na * FT * 2; idx += kNTP) {
                const int h = idx & 1, fl = (idx >> 1) & (FT - 1), a = idx >> (LFT + 1);
                const unsigned off = (unsigned)(a * AS + fl * kMatBytes + h * 32);
                const unsigned char *e1 = (p.same_dde ? e2 : am) + off;
                const unsigned char *bm = bsm + fl * kMatBytes;
                const Cd b0 = lds_c(bm), b1 = lds_c(bm + 16), b2 = lds_c(bm + 32), b3 = lds_c(bm + 48);
                Cd x0, x1;
                if constexpr (SAMPLE) {
                    // row h of the beam Jones of antenna a at channel fl from its two frequency planes:
                    // the arithmetic of beam_cube_dde_planes_kernel's stage 2 (afr_beam.cu), i.e. the
                    // reference's interpolation and amplitude renormalisation (fast_beam_cubes.py:169-238)
                    const double2 *pr = reinterpret_cast<const double2 *>(pl_of(st) + (size_t)((a << LFT) + fl) * kPlaneStride) + h;
                    const double wlo = fw[fl], whi = 1.0 - wlo;  // fd holds the weight; 1 - w as the beam kernel forms it
                    // the row's two correlations are adjacent: one 16-byte load per (plane, re / im / |.|)
                    const double2 rlo = pr[0], ilo = pr[2], alo = pr[4], rhi = pr[6], ihi = pr[8], ahi = pr[10];
                    Cd xs[2];
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        const double csr = fma(wlo, cc ? rlo.y : rlo.x, whi * (cc ? rhi.y : rhi.x));
                        const double csi = fma(wlo, cc ? ilo.y : ilo.x, whi * (cc ? ihi.y : ihi.x));
                        const double asum = fma(wlo, cc ? alo.y : alo.x, whi * (cc ? ahi.y : ahi.x));
                        const double tt = fma(csr, csr, csi * csi);
                        double kk;
                        if (tt > 1e-290 && tt < 1e290) {
                            kk = asum * rsqrt(tt);
                        } else {
                            const double div = hypot(csr, csi);
                            kk = (div == 0.0) ? asum : asum / div;
                        }
                        xs[cc] = {csr * kk, csi * kk};
                    }
                    x0 = xs[0], x1 = xs[1];
                    if (p.feed != nullptr) {
                        // dde = beam . L[t,a] (rime/examples/predict.py:469-472): row h of E times the feed rotation
                        const double2 *L = reinterpret_cast<const double2 *>(p.feed) + ((long long)t * p.nant + a) * 4;
                        const double2 l00 = L[0], l01 = L[1], l10 = L[2], l11 = L[3];
                        const Cd y0 = cadd_(cmul_(x0, Cd{l00.x, l00.y}), cmul_(x1, Cd{l10.x, l10.y}));
                        const Cd y1 = cadd_(cmul_(x0, Cd{l01.x, l01.y}), cmul_(x1, Cd{l11.x, l11.y}));
                        x0 = y0, x1 = y1;
                    }
                } else {
                    x0 = lds_c(e1), x1 = lds_c(e1 + 16);
                }
                Cd m0 = cadd_(cmul_(x0, b0), cmul_(x1, b2));
                Cd m1 = cadd_(cmul_(x0, b1), cmul_(x1, b3));
                if (ANT) {
                    // antenna phasor k_a(f) = exp(i psi_a nu_f), psi_a from the antenna coordinates
                    const double psi = __dmul_rn(
                        p.cst, phase_dot(sl, sm, sn, ant_t[3 * a], ant_t[3 * a + 1], ant_t[3 * a + 2], false));
                    const C2<double> kk = cis_fast(__dmul_rn(psi, fq[fl]));
                    const Cd k = {kk.re, kk.im};
                    m0 = cmul_(k, m0), m1 = cmul_(k, m1);
                    if (p.same_dde) {
                        sts_c(e2 + off, cmul_(k, x0));
                        sts_c(e2 + off + 16, cmul_(k, x1));
                    } else {
                        sts_c(e2 + off, cmul_(k, lds_c(e2 + off)));
                        sts_c(e2 + off + 16, cmul_(k, lds_c(e2 + off + 16)));
                    }
                }
                sts_c(am + off, m0);
                sts_c(am + off + 16, m1);
            }
            mbar_arrive(&bars[st]);  // full: A (and, antenna mode, the scaled E2) of source s are visible
            // ---- next copies: source s + kNS - 1 goes into the stage source s - 1 used
            if (s + kNS - 1 < nsrc) {
                if (s >= 1) mbar_wait(&bars[kNS + (int)((s - 1) % kNS)], (unsigned)(((s - 1) / kNS) & 1));
                issue(s + kNS - 1);
            }
        }
        return;
    }

    // ================================= CONSUMERS =================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;\n");
    // Row k of this thread: perm[rbeg + k * 512 + tid].  perm orders the rows of a timestep by
    // (antenna1 / 4, antenna2 / 8) tiles, so the 32 lanes of a warp read <= 4 distinct A and
    // <= 8 distinct E matrices in distinct bank groups: every LDS.128 is one wavefront
    // instead of four (shared-memory bandwidth was the limit: l1tex 83 % busy before).
    constexpr bool kPairs = ANT && RPT == 4;
    unsigned offs[RPT];  // shared-memory offsets of (antenna1 | antenna2 << 16)
    int rowid[RPT];
    double ru = 0.0, rv = 0.0, rw = 0.0;  // ROW mode (RPT == 1): uvw of this thread's row
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
        // RPT == 4 (antenna mode): rows 2pp, 2pp+1 of a thread are CONSECUTIVE in perm, which the
        // pair ordering of launch_row_tile_order makes two baselines of the same antenna 1
        const long long ri = kPairs ? rbeg + 2 * ((k >> 1) * kConsThreads + tid) + (k & 1)
                                    : rbeg + k * kConsThreads + tid;
        offs[k] = 0;
        rowid[k] = -1;
        if (ri < rend) {
            const int r = p.perm[ri];
            rowid[k] = r;
            offs[k] = (unsigned)(p.ant1[r] * AS) | ((unsigned)(p.ant2[r] * AS) << 16);
            if (!ANT) {
                ru = p.uvw[3 * (long long)r];
                rv = p.uvw[3 * (long long)r + 1];
                rw = p.uvw[3 * (long long)r + 2];
            }
        }
    }

    Cd acc[RPT][FT][4];
#pragma unroll
    for (int k = 0; k < RPT; ++k)
#pragma unroll
        for (int j = 0; j < FT; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[k][j][c] = {0.0, 0.0};

    for (long long s = 0; s < nsrc; ++s) {
        const int st = (int)(s % kNS);
        // ROW mode: this thread's own phasor anchors, computed while the producers prepare the
        // stage (the wait for it comes after);
        // the phase argument is rounded exactly as the reference rounds it (phase.py:49-53)
        Cd z = {1.0, 0.0}, zp = {1.0, 0.0}, d = {1.0, 0.0};
        double c2 = 2.0;
        if (!ANT) {
            const double phi = __dmul_rn(
                p.cst, phase_dot(p.lmn[3 * s], p.lmn[3 * s + 1], p.lmn[3 * s + 2], ru, rv, rw, false));
            if (EXACT) {
                d.re = phi;
            } else {
                const C2<double> zz = cis_fast(__dmul_rn(phi, nu0)), dd = cis_fast(__dmul_rn(phi, dnu));
                z = {zz.re, zz.im};
                d = {dd.re, dd.im};
                c2 = d.re + d.re;
            }
        }
        mbar_wait(&bars[st], (unsigned)((s / kNS) & 1));
        if constexpr (kPairs) {
            // Two baselines sharing antenna 1: A_{a1} stays in registers (4 LDS.128 per pair), each
            // E_{a2} streams through as two half-matrices (2 x 2 LDS.128): 12 instead of 16
            // LDS.128 per two terms.  Shared memory delivers 128 B per clock per SM however many
            // lanes share an address, and it -- not the FP64 pipe -- bounds this loop.
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
                const unsigned char *am = a_of(st) + (offs[2 * pp] & 0xFFFFu);
                Cd x0 = lds_c(am), x1 = lds_c(am + 16), x2 = lds_c(am + 32), x3 = lds_c(am + 48);
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int k = 2 * pp + r;
                    if (r == 1 && (offs[k] & 0xFFFFu) != (offs[k - 1] & 0xFFFFu)) {
                        // a pair straddling two antenna-1 runs (at most one per run)
                        const unsigned char *am1 = a_of(st) + (offs[k] & 0xFFFFu);
                        x0 = lds_c(am1), x1 = lds_c(am1 + 16), x2 = lds_c(am1 + 32), x3 = lds_c(am1 + 48);
                    }
                    const unsigned char *e2 = e2_of(st) + (offs[k] >> 16);
#pragma unroll
                    for (int c = 0; c < 2; ++c) {  // column c of A E^H needs row c of E
                        const Cd q0 = lds_c(e2 + 32 * c), q1 = lds_c(e2 + 32 * c + 16);
                        cfmac_(acc[k][0][c], x0, q0);
                        cfmac_(acc[k][0][2 + c], x2, q0);
                        cfmac_(acc[k][0][c], x1, q1);
                        cfmac_(acc[k][0][2 + c], x3, q1);
                    }
                }
            }
        } else {
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const unsigned char *e2 = e2_of(st) + (offs[k] >> 16);
            const unsigned char *am = a_of(st) + (offs[k] & 0xFFFFu);
#pragma unroll
            for (int j = 0; j < FT; ++j) {
                if (!ANT && EXACT) {
                    const C2<double> zz = cis_fast(__dmul_rn(d.re, fq[j]));
                    z = {zz.re, zz.im};
                }
                const Cd q0 = lds_c(e2 + j * kMatBytes), q1 = lds_c(e2 + j * kMatBytes + 16);
                const Cd q2 = lds_c(e2 + j * kMatBytes + 32), q3 = lds_c(e2 + j * kMatBytes + 48);
#pragma unroll
                for (int h = 0; h < 2; ++h) {  // one output row of A * E^H at a time
                    const Cd x0 = lds_c(am + j * kMatBytes + 32 * h);
                    const Cd x1 = lds_c(am + j * kMatBytes + 32 * h + 16);
                    if (ANT) {
                        // 16 DFMA per output row, accumulated in place (32 per term; the
                        // mul / fma / add form was 48 FP64 instructions per term)
                        cfmac_(acc[k][j][2 * h], x0, q0);
                        cfmac_(acc[k][j][2 * h + 1], x0, q2);
                        cfmac_(acc[k][j][2 * h], x1, q1);
                        cfmac_(acc[k][j][2 * h + 1], x1, q3);
                    } else {
                        Cd m0 = {x0.re * q0.re, x0.im * q0.re}, m1 = {x0.re * q2.re, x0.im * q2.re};
                        m0.re = fma(x0.im, q0.im, m0.re), m0.im = fma(-x0.re, q0.im, m0.im);
                        m1.re = fma(x0.im, q2.im, m1.re), m1.im = fma(-x0.re, q2.im, m1.im);
                        cfmac_(m0, x1, q1);
                        cfmac_(m1, x1, q3);
                        cfma_(acc[k][j][2 * h], z, m0);
                        cfma_(acc[k][j][2 * h + 1], z, m1);
                    }
                }
                if (!ANT && !EXACT && j + 1 < FT) {
                    // three-term recurrence z_{j+1} = 2 Re(d) z_j - z_{j-1} (first step: z * d)
                    Cd zn;
                    if (j == 0) {
                        zn = cmul_(z, d);
                    } else {
                        zn.re = fma(c2, z.re, -zp.re);
                        zn.im = fma(c2, z.im, -zp.im);
                    }
                    zp = z;
                    z = zn;
                }
            }
        }
        }
        __syncwarp();
        if (p.arrive_all || lane == 0) mbar_arrive(&bars[kNS + st]);
    }

#pragma unroll
    for (int k = 0; k < RPT; ++k) {
        const long long r = rowid[k];
        if (r >= 0) {
#pragma unroll
            for (int j = 0; j < FT; ++j) {
                const int f = f0 + j;
                if (f < p.nchan) {
                    double2 *o = reinterpret_cast<double2 *>(p.out + (r * p.nchan + f) * 8);
#pragma unroll
                    for (int c = 0; c < 4; ++c) o[c] = make_double2(acc[k][j][c].re, acc[k][j][c].im);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Antenna decomposition of the baseline uvw of each timestep (ANT mode).
//
// The uvw of a row are, in every Measurement Set, the difference of two per-antenna
// coordinates: uvw[r] = U[a1] - U[a2] (or the opposite sign: the same decomposition with U
// negated).  If that holds for the caller's arrays the phasor factorises per antenna.  One CTA
// per timestep: the reference antenna is the antenna1 of the timestep's first row, U[ref] = 0,
// U[a] follows from the rows that join `a` to the reference; then EVERY row is checked.
// The mode is only allowed when the phase it produces provably stays within 5e-11 rad of the
// phase the reference rounds (half the 1e-10 parity gate):
//     |cst| nu_max L (|uvw[r] - (U[a1] - U[a2])| + 9 eps max(|U[a1]|, |U[a2]|)) <= 5e-11,
// L = max_s (|l| + |m| + |n|): the first term is the decomposition residual, the second the
// rounding of the two per-antenna phase arguments plus the reference's own.  Wide fields or
// very long baselines therefore stay in ROW mode.  Any antenna not joined to the reference,
// or any row that fails, clears ok[0].
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) antenna_uvw_kernel(const double *uvw, const int32_t *ant1,
                                                          const int32_t *ant2, const int32_t *row_start,
                                                          long long nant, const double *lmn,
                                                          long long nsrc, const double *freq,
                                                          long long nchan, double cst, double *ant_uvw,
                                                          int *ok) {
    extern __shared__ unsigned char sm_raw[];
    int *have = reinterpret_cast<int *>(sm_raw);  // [nant]
    __shared__ double red[2][256];
    const int t = blockIdx.x;
    const long long rb = row_start[t], re = row_start[t + 1];
    double *U = ant_uvw + (long long)t * nant * 3;
    for (int a = threadIdx.x; a < nant; a += blockDim.x) {
        have[a] = 0;
        U[3 * a] = U[3 * a + 1] = U[3 * a + 2] = 0.0;
    }
    double L = 0.0, nu = 0.0;
    for (long long s = threadIdx.x; s < nsrc; s += blockDim.x)
        L = fmax(L, fabs(lmn[3 * s]) + fabs(lmn[3 * s + 1]) + fabs(lmn[3 * s + 2]));
    for (long long f = threadIdx.x; f < nchan; f += blockDim.x) nu = fmax(nu, fabs(freq[f]));
    red[0][threadIdx.x] = L;
    red[1][threadIdx.x] = nu;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) {
            red[0][threadIdx.x] = fmax(red[0][threadIdx.x], red[0][threadIdx.x + w]);
            red[1][threadIdx.x] = fmax(red[1][threadIdx.x], red[1][threadIdx.x + w]);
        }
        __syncthreads();
    }
    const double rad_per_metre = fabs(cst) * red[1][0] * red[0][0];
    if (rb >= re) return;
    const int ref = ant1[rb];
    if (threadIdx.x == 0) have[ref] = 1;
    __syncthreads();
    for (long long r = rb + threadIdx.x; r < re; r += blockDim.x) {
        const int p1 = ant1[r], p2 = ant2[r];
        if (p1 == ref && p2 != ref) {
            U[3 * p2] = -uvw[3 * r];
            U[3 * p2 + 1] = -uvw[3 * r + 1];
            U[3 * p2 + 2] = -uvw[3 * r + 2];
            have[p2] = 1;
        } else if (p2 == ref && p1 != ref) {
            U[3 * p1] = uvw[3 * r];
            U[3 * p1 + 1] = uvw[3 * r + 1];
            U[3 * p1 + 2] = uvw[3 * r + 2];
            have[p1] = 1;
        }
    }
    __syncthreads();
    bool good = true;
    for (long long r = rb + threadIdx.x; r < re; r += blockDim.x) {
        const int p1 = ant1[r], p2 = ant2[r];
        if (!have[p1] || !have[p2]) {
            good = false;
            break;
        }
        double worst = 0.0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double u1 = U[3 * p1 + c], u2 = U[3 * p2 + c], got = uvw[3 * r + c];
            worst = fmax(worst, fabs(got - (u1 - u2)) + 9.0 * 1.1102230246251565e-16 * fmax(fabs(u1), fabs(u2)));
        }
        if (!(rad_per_metre * worst <= 5e-11)) good = false;  // also catches NaN
    }
    if (!good) atomicAnd(ok, 0);
}

// sort key of a row: time | antenna tile (a1/4, a2/8) | position inside the tile.
// pair_order (antenna mode, 2048-row tiles): time | a2/16 | a1/4 | a1 % 4 | a2 % 8 | (a2/8) % 2 --
// tiles of 4 x 16 antennas = 64 rows = one warp of row PAIRS: consecutive rows are (a1, q) and
// (a1, q + 8), so a thread's two consecutive rows share antenna 1 (except where a ragged tile
// shifts the pairing across an antenna-1 boundary), and a warp still touches only ~4 distinct A and
// ~8 + 8 distinct E matrices in distinct bank groups (shared-memory wavefronts go with the
// distinct data a warp reads: a run-ordered pairing with 32 distinct E per load was measured
// SLOWER than no pairing, 214 vs 228 Gterms/s, with 5x the bank conflicts).
__global__ void row_keys_kernel(const int32_t *time_index, const int32_t *ant1, const int32_t *ant2,
                                long long nrow, int pair_order, unsigned long long *keys, int32_t *rows) {
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (r >= nrow) return;
    const unsigned a1 = (unsigned)ant1[r] & 1023u, a2 = (unsigned)ant2[r] & 1023u;
    const unsigned long long tkey = (unsigned long long)(unsigned)time_index[r] << 20;
    if (pair_order)
        keys[r] = tkey | ((a2 >> 4) << 14) | ((a1 >> 2) << 6) | ((a1 & 3u) << 4) | ((a2 & 7u) << 1) |
                  ((a2 >> 3) & 1u);
    else
        keys[r] = tkey | ((a1 >> 2) << 12) | ((a2 >> 3) << 5) | ((a1 & 3u) << 3) | (a2 & 7u);
    rows[r] = (int32_t)r;
}

}  // namespace

// perm (nrow): the rows ordered by (time, antenna tile); rows of one timestep stay contiguous
int launch_row_tile_order(const int32_t *time_index, const int32_t *ant1, const int32_t *ant2,
                          int64_t nrow, int64_t ntime, bool pair_order, int32_t *perm,
                          cudaStream_t stream) {
    if (nrow <= 0) return 0;
    Scratch keys_in, keys_out, rows_in, tmp;
    AFR_CUDA_OK(keys_in.alloc(sizeof(unsigned long long) * (size_t)nrow, stream));
    AFR_CUDA_OK(keys_out.alloc(sizeof(unsigned long long) * (size_t)nrow, stream));
    AFR_CUDA_OK(rows_in.alloc(sizeof(int32_t) * (size_t)nrow, stream));
    row_keys_kernel<<<(unsigned)((nrow + 255) / 256), 256, 0, stream>>>(
        time_index, ant1, ant2, nrow, pair_order ? 1 : 0, (unsigned long long *)keys_in.ptr,
        (int32_t *)rows_in.ptr);
    AFR_LAUNCH_OK();
    int tbits = 1;
    while ((1LL << tbits) < ntime && tbits < 43) ++tbits;
    size_t tmp_bytes = 0;
    AFR_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const unsigned long long *)keys_in.ptr,
                                                (unsigned long long *)keys_out.ptr,
                                                (const int32_t *)rows_in.ptr, perm, (int)nrow, 0,
                                                20 + tbits, stream));
    AFR_CUDA_OK(tmp.alloc(tmp_bytes, stream));
    AFR_CUDA_OK(cub::DeviceRadixSort::SortPairs(tmp.ptr, tmp_bytes, (const unsigned long long *)keys_in.ptr,
                                                (unsigned long long *)keys_out.ptr,
                                                (const int32_t *)rows_in.ptr, perm, (int)nrow, 0,
                                                20 + tbits, stream));
    note_launch(3);  // cub's histogram / onesweep kernels
    return 0;
}

size_t dde_ws_smem_bytes(int64_t nant, int ft, bool ant, bool sample) {
    return kNS * stage_bytes((int)nant, ft, ant, sample) + 3 * kNS * sizeof(uint64_t) + 12 * sizeof(double);
}

int launch_antenna_uvw(const double *uvw, const int32_t *ant1, const int32_t *ant2,
                       const int32_t *row_start, int64_t ntime, int64_t nant, const double *lmn,
                       int64_t nsrc, const double *freq, int64_t nchan, double cst, double *ant_uvw,
                       int *ok, cudaStream_t stream) {
    if (ntime <= 0) return 0;
    antenna_uvw_kernel<<<(unsigned)ntime, 256, (size_t)nant * sizeof(int), stream>>>(
        uvw, ant1, ant2, row_start, nant, lmn, nsrc, freq, nchan, cst, ant_uvw, ok);
    AFR_LAUNCH_OK();
    return 0;
}

// channels per CTA of the 512-row tile: the most of 4, 2, 1 whose three-stage antenna tile fits
// in shared memory (4 up to 138 antennas, 2 up to 258, 1 up to 469); 0 = none fits
int dde_ws_row_tile_channels(int64_t nant) {
    for (int ft = 4; ft >= 1; ft /= 2)
        if (dde_ws_smem_bytes(nant, ft, false) <= 220 * 1024) return ft;
    return 0;
}

int launch_fused_dde_ws(const DdeWsParams &p, int max_rows_per_time, bool exact, bool ant_mode,
                        cudaStream_t stream) {
    // antenna mode with more than one 512-row tile per timestep: 2048 rows x 1 channel per CTA
    const bool wide_rows = ant_mode && max_rows_per_time > kConsThreads;
    const bool sample = p.planes != nullptr;
    AFR_REQUIRE(!sample || (ant_mode && p.same_dde && !exact), "in-kernel beam sampling needs antenna mode");
    int ft = wide_rows ? 1 : dde_ws_row_tile_channels(p.nant);
    const int rpt = wide_rows ? 4 : 1;
    while (sample && ft > 1 && dde_ws_smem_bytes(p.nant, ft, true, true) > 220 * 1024) ft /= 2;
    AFR_REQUIRE(ft > 0 && dde_ws_smem_bytes(p.nant, ft, ant_mode, sample) <= 220 * 1024,
                "afr_predict_fused: antenna tile does not fit in shared memory");
    const size_t smem = dde_ws_smem_bytes(p.nant, ft, ant_mode, sample);
    const int rows = kConsThreads * rpt;
    dim3 grid((unsigned)((max_rows_per_time + rows - 1) / rows), (unsigned)p.ntime,
              (unsigned)((p.nchan + ft - 1) / ft));
    AFR_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "afr_predict_fused: grid too large");
    AFR_REQUIRE((size_t)p.nant * ant_stride(ft) < 65536, "afr_predict_fused: too many antennas");
    const int threads = (kConsWarps + kProdWarps) * 32;
    DdeWsParams q = p;
    q.arrive_all = (getenv("AFR_SANITIZE") && atoi(getenv("AFR_SANITIZE")) != 0) ? 1 : 0;
    auto go = [&](auto kern) -> int {
        cudaFuncAttributes attr;
        AFR_CUDA_OK(cudaFuncGetAttributes(&attr, kern));
        // setmaxnreg moves registers inside the launch-time pool: it must hold 512 x 104 + 128 x 64
        AFR_REQUIRE(threads * attr.numRegs >= kConsWarps * 32 * 104 + kProdWarps * 32 * 64,
                    "fused_dde_ws: launch-time register pool too small for setmaxnreg");
        AFR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, threads, smem, stream>>>(q);
        return 0;
    };
    int rc;
    if (sample && wide_rows)
        rc = go(fused_dde_ws_kernel<false, true, 1, 4, true>);
    else if (sample)
        rc = ft == 4 ? go(fused_dde_ws_kernel<false, true, 4, 1, true>)
                     : (ft == 2 ? go(fused_dde_ws_kernel<false, true, 2, 1, true>)
                                : go(fused_dde_ws_kernel<false, true, 1, 1, true>));
    else if (wide_rows)  // antenna phasors are evaluated per channel: any frequency array
        rc = go(fused_dde_ws_kernel<false, true, 1, 4>);
    else if (ant_mode)
        rc = ft == 4 ? go(fused_dde_ws_kernel<false, true, 4, 1>)
                     : (ft == 2 ? go(fused_dde_ws_kernel<false, true, 2, 1>)
                                : go(fused_dde_ws_kernel<false, true, 1, 1>));
    else if (exact)
        rc = ft == 4 ? go(fused_dde_ws_kernel<true, false, 4, 1>)
                     : (ft == 2 ? go(fused_dde_ws_kernel<true, false, 2, 1>)
                                : go(fused_dde_ws_kernel<true, false, 1, 1>));
    else
        rc = ft == 4 ? go(fused_dde_ws_kernel<false, false, 4, 1>)
                     : (ft == 2 ? go(fused_dde_ws_kernel<false, false, 2, 1>)
                                : go(fused_dde_ws_kernel<false, false, 1, 1>));
    if (rc) return rc;
    AFR_LAUNCH_OK();
    return 0;
}

}  // namespace afr
