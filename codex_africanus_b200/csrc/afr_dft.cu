// Phasor-stream kernels: im_to_vis, vis_to_im (and the point-source part of the
// fused predict) for sm_100a.
//
// Mathematical object (africanus/dft/kernels.py:23-69, 83-148):
//     acc[x, f, c] += exp(i * cst * (X[x] . Y[y]) * nu_f) * W[y, f, c]      summed over y
// im_to_vis : owners x = rows (u,v,w), streamed y = sources (l,m,n), W = image, complex acc
// vis_to_im : owners x = sources,      streamed y = rows,            W = vis,   real part only
//
// B200 mapping (the path is FP64-pipe bound, bytes/term << 1, no GEMM is pretended):
//  * one lane owns one x (32 owners per warp), one warp owns a run of CH channels, so a
//    thread keeps CH*ncorr accumulators in registers for the whole y loop; a CTA is NW warps
//    = (owner groups) x (channel runs);
//  * y is streamed in tiles of yt items through DOUBLE-BUFFERED shared memory with one
//    barrier per tile: the W tile and the y coordinates arrive by cp.async (16-byte
//    coalesced copies, zero-filled at the edges) issued one tile ahead, and the per-(x,y)
//    ANCHORS for tile t+1 are computed by the same warps right after they consume tile t;
//  * inside a channel run the phasor advances by ONE complex rotation per channel
//    (2 DMUL + 2 DFMA) instead of a sincos; every run restarts from an anchor
//    exp(i*phi*nu_f0), so the recurrence never runs longer than CH <= 32 steps.  Per (x,y)
//    pair the anchors cost 2 sincos (first run, channel step d), log2(CH) complex squarings
//    (run step D = d^CH) and one rotation per further run;
//  * the phase argument phi = cst*(l*u + m*v + n*w) is formed in FP64 in the reference's
//    operation order with explicitly rounded ops (no FMA contraction), and the first anchor
//    phase is fl(phi*nu_f0) exactly as the reference computes it;
//  * non-equispaced channels (exact mode) take one sincos per term instead;
//  * two y's are processed per inner iteration so each thread has two independent
//    rotation chains in flight (measured DFMA latency 8.4 cycles, issue 1 per 2 cycles/SMSP).
#include <algorithm>
#include <cstdlib>

#include "afr_dft.cuh"

namespace afr {

namespace {

constexpr int kMaxChunks = 8;  // cp.async granules of the W tile per thread

struct DftParams {
    const double *xc;      // (nx,3) owner coordinates
    const double *yc;      // (ny,3) streamed coordinates
    const void *w;         // (ny,nchan,wstride) real or complex, in the accumulator precision
    const uint8_t *anyflag;  // (ny,nchan) 1 = drop the sample, or nullptr
    const double *freq;    // (nchan,)
    void *out;             // (nsplit,nx,nchan,wstride) in the accumulator type
    double cst;
    long long nx, ny;
    long long ysplit;            // y items per grid.z slice (multiple of yt)
    long long out_split_stride;  // scalars between slices
    int nchan;
    int wstride;  // correlations in W / out
    int coff;     // first correlation handled by this launch
    int nck;      // channel runs per CTA (power of two <= NW/2)
    int yt;       // y items per tile (even, <= 8)
    int f32dot;
    int fast;         // 1: W rows are contiguous and copied by cp.async granules
    int granule;      // 4, 8 or 16 bytes
    int row_chunks_log2;  // log2(granules per W tile row)
    int bulk;             // warp-specialised kernel: W tile by TMA bulk copies
    int arrive_all;       // every consumer lane arrives on the "empty" mbarrier (AFR_SANITIZE=1)
    int one;              // 1 (opaque to the compiler: trip count of the scalar-burst blocks, see the MMA consumers)
    int ablate;           // diagnostics (AFR_POINT_MMA_ABLATE, wrong results): 1 no anchor maths, 2 no consumer recurrence
};

__device__ __forceinline__ void cp_async(unsigned dst, const void *src, int granule, int src_bytes) {
    if (granule == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
    else if (granule == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

template <int N>
__device__ __forceinline__ void load_vec(const double *src, double (&dst)[N]) {
    static_assert(N % 2 == 0, "16-byte granularity");
    const double2 *s = reinterpret_cast<const double2 *>(src);
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        double2 v = s[i];
        dst[2 * i] = v.x;
        dst[2 * i + 1] = v.y;
    }
}

template <int N>
__device__ __forceinline__ void load_vec(const float *src, float (&dst)[N]) {
    static_assert(N % 4 == 0, "16-byte granularity");
    const float4 *s = reinterpret_cast<const float4 *>(src);
#pragma unroll
    for (int i = 0; i < N / 4; ++i) {
        float4 v = s[i];
        dst[4 * i] = v.x;
        dst[4 * i + 1] = v.y;
        dst[4 * i + 2] = v.z;
        dst[4 * i + 3] = v.w;
    }
}

__device__ __noinline__ C2<double> cis_noinline(double p) { return cis_fast(p); }

// acc (+)= z * w for the NCORR correlations of one channel
template <int NCORR, bool WC, bool ADJ, typename ACC>
__device__ __forceinline__ void accumulate(ACC (&are)[NCORR], ACC (&aim)[ADJ ? 1 : NCORR],
                                           const C2<ACC> z, const ACC *wv) {
#pragma unroll
    for (int c = 0; c < NCORR; ++c) {
        if (WC) {
            const ACC wr = wv[2 * c], wi = wv[2 * c + 1];
            are[c] = fma(z.re, wr, are[c]);
            are[c] = fma(-z.im, wi, are[c]);
            if (!ADJ) {
                aim[c] = fma(z.re, wi, aim[c]);
                aim[c] = fma(z.im, wr, aim[c]);
            }
        } else {
            const ACC wr = wv[c];
            are[c] = fma(z.re, wr, are[c]);
            if (!ADJ) aim[c] = fma(z.im, wr, aim[c]);
        }
    }
}

// Anchors of the nck channel runs of one (x, y) pair: a_k = a * D^k, written to
// anch[k * stride].  The chain across runs uses the same three-term form as the channels
// inside a run (a_{k+1} = 2 Re(D) a_k - a_{k-1}): half the FP64 instructions and half the
// dependent latency of a complex multiply per run; <= 16 steps, error < 3e-14.
template <typename ACC>
__device__ __forceinline__ void store_run_anchors(C2<ACC> *anch, int nck, int stride, C2<double> a,
                                                  const C2<double> D) {
    auto put = [&](const C2<double> v) {
        C2<ACC> t;
        t.re = (ACC)v.re;
        t.im = (ACC)v.im;
        *anch = t;
        anch += stride;
    };
    put(a);
    if (nck == 1) return;  // nck is a power of two
    C2<double> a1 = cmul(a, D);
    put(a1);
    const double c2 = D.re + D.re;
#pragma unroll 2
    for (int k = 2; k < nck; k += 2) {
        C2<double> a2, a3;
        a2.re = fma(c2, a1.re, -a.re);
        a2.im = fma(c2, a1.im, -a.im);
        a3.re = fma(c2, a2.re, -a1.re);
        a3.im = fma(c2, a2.im, -a1.im);
        put(a2);
        put(a3);
        a = a2;
        a1 = a3;
    }
}

// FP32 runs longer than 8 channels are cut into sub-runs (see consume_run)
constexpr int kSubRun = 16;  // channels per FP32 sub-run
template <typename ACC, int CH>
constexpr bool kSubRuns = sizeof(ACC) == 4 && (CH > kSubRun);

// exp(i*phi*nu0), exp(i*phi*dnu) and the run step D = d^CH (repeated squaring) of one pair;
// d8 = d^8 (an intermediate of the squaring) for the FP32 sub-runs of consume_run
template <int CH>
__device__ __forceinline__ void pair_anchors(double phi, double nu0, double dnu, bool need_D,
                                             C2<double> &a, C2<double> &d, C2<double> &D,
                                             C2<double> *d8 = nullptr) {
    a = cis_fast(__dmul_rn(phi, nu0));
    d = cis_fast(__dmul_rn(phi, dnu));
    D = d;
    if (need_D || d8 != nullptr) {
#pragma unroll
        for (int q = 1; q < CH; q *= 2) {
            if (q == kSubRun && d8 != nullptr) *d8 = D;
            if (q >= kSubRun && !need_D) break;
            const double re = D.re * D.re - D.im * D.im;
            D.im = 2.0 * D.re * D.im;
            D.re = re;
        }
    }
}

// One channel run of one (x, y) pair: acc[j] (+)= z_j * W[j], z_j = z0 * d^j, j < CH.
//
// FP64: the phasor advances by the THREE-TERM form of the rotation recurrence,
//     z_{j+1} = (2 Re d) z_j - z_{j-1}          (2 DFMA per channel instead of 2 DMUL + 2 DFMA),
// started from the two anchors z_0 (exact sincos) and z_1 = z_0 * d, and restarted at every
// run (CH <= 32 channels).  Rounding of the coefficient perturbs z_j by at most
// eps * (|U'_{j-1}| + |U'_{j-2}|) <= (2/3) j^3 eps (U = Chebyshev polynomials of the 2nd
// kind), injected roundings by j^2/2 eps; measured over 2e7 steps incl. d -> 0 and d -> pi
// the worst error after 32 channels is 1.3e-13 (oracle/../tests: uniform-vs-exact goldens),
// three orders below the 1e-10 gate.  In FP32 a 32-channel three-term run would cost ~1e3 eps_32 =
// 6e-5, above the 1e-5 gate: FP32 runs are cut into sub-runs of 8 (below).
// FP32 (complex64 / float32 outputs): the run is cut into SUB-RUNS of 8 channels.  Inside a sub-run
// the phasor advances by the three-term form in FP32 (2 FFMA per channel sharing the coefficient
// operand, where the plain rotation is 2 FMUL + 2 FFMA): over 7 steps its error stays below
// ~j^2 eps_32 = 3e-6 (worst case, step angle -> 0; measured rms an order lower), inside the 1e-5
// gate, where 31 steps would reach 6e-5.  Sub-run starts are plain FP32 rotations of the run's
// FP64-computed anchor by d8 = d^8 (formed in FP64): 3 roundings at most.  6 -> 4.5 FP32 instructions
// per term.
template <int NCORR, bool WC, bool ADJ, typename ACC, int CH, int G>
__device__ __forceinline__ void consume_run(ACC (&are)[CH][NCORR], ACC (&aim)[CH][ADJ ? 1 : NCORR],
                                            C2<ACC> z, const C2<ACC> d, const ACC *wrow,
                                            const C2<ACC> d8 = C2<ACC>{ACC(1), ACC(0)}) {
    constexpr int NV = NCORR * (WC ? 2 : 1);
    constexpr bool kF32 = sizeof(ACC) == 4;
    constexpr bool kThreeTerm = CH > 2;
    constexpr int SUB = kF32 ? kSubRun : CH;  // channels between restarts of the three-term recurrence
    C2<ACC> zp = z;   // z_{j-1}
    C2<ACC> zs = z;   // start of the current sub-run
    const ACC c2 = d.re + d.re;
#pragma unroll
    for (int j = 0; j < CH; j += G) {
        ACC wv[G * NV];
        load_vec<G * NV>(wrow + j * NV, wv);
#pragma unroll
        for (int g = 0; g < G; ++g) {
            accumulate<NCORR, WC, ADJ, ACC>(are[j + g], aim[j + g], z, wv + g * NV);
            if (j + g + 1 < CH) {
                const int s = (j + g) % SUB;  // position inside the sub-run
                if (s == SUB - 1) {           // next channel starts a sub-run: rotate its start by d^8
                    zs = cmul(zs, d8);
                    z = zs;
                } else if (!kThreeTerm || s == 0) {
                    const C2<ACC> zn = cmul(z, d);
                    zp = z;
                    z = zn;
                } else {
                    C2<ACC> zn;
                    zn.re = fma(c2, z.re, -zp.re);
                    zn.im = fma(c2, z.im, -zp.im);
                    zp = z;
                    z = zn;
                }
            }
        }
    }
}

template <int NCORR, bool WC, bool ADJ, typename ACC, int CH, int NW, bool EXACT>
__global__ void __launch_bounds__(NW * 32, 1) phasor_stream_kernel(const DftParams p) {
    constexpr int NT = NW * 32;
    constexpr int NV = NCORR * (WC ? 2 : 1);  // W scalars per channel
    constexpr int G = (NV * (int)sizeof(ACC) >= 16) ? 1 : 16 / (NV * (int)sizeof(ACC));  // channels per 16-byte W load
    constexpr int SZ = (int)sizeof(ACC);
    static_assert(CH % G == 0 && (G * NV * SZ) % 16 == 0, "tiling");
    static_assert((CH & (CH - 1)) == 0, "CH must be a power of two");
    using CA = C2<ACC>;

    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nck = p.nck, yt = p.yt;
    const int xgw = (NW / nck) * 32;  // owners per CTA
    const int ft = nck * CH;          // channels per CTA
    const int cta_f0 = blockIdx.y * ft;
    const long long cta_x0 = (long long)blockIdx.x * xgw;

    // ---- shared-memory carve-up: two {anchor, step, W} buffers + 3 y-coordinate slots
    constexpr bool kSub = kSubRuns<ACC, CH>;  // FP32 sub-runs: the pair's d^8 rides along with d
    constexpr int DS = kSub ? 2 : 1;
    const size_t anch_elems = (size_t)yt * nck * xgw;  // CA, or double phi[yt][xgw] (exact)
    const size_t dstp_elems = (size_t)yt * xgw * DS;
    const size_t w_elems = (size_t)yt * ft * NV;
    const size_t buf_bytes = (anch_elems + dstp_elems) * sizeof(CA) + w_elems * SZ;
    double *ycs = reinterpret_cast<double *>(smem_raw + 2 * buf_bytes);  // [3][yt*3]
    double *fq = ycs + 3 * (size_t)yt * 3;                               // [ft] (exact mode)
    auto anch_of = [&](int b) { return reinterpret_cast<CA *>(smem_raw + (size_t)b * buf_bytes); };
    auto dstp_of = [&](int b) { return anch_of(b) + anch_elems; };
    auto w_of = [&](int b) { return reinterpret_cast<ACC *>(dstp_of(b) + dstp_elems); };

    // consumer role: lane -> owner, warp -> (owner group, channel run)
    const int ck = warp % nck;
    const int x_local = (warp / nck) * 32 + lane;
    const long long x = cta_x0 + x_local;
    const int fo = ck * CH;

    // producer role: fixed owner per thread, y strided
    const int px_local = tid % xgw;
    const int py0 = tid / xgw;
    const int pystep = NT / xgw > 0 ? NT / xgw : 1;
    double px0, px1, px2;
    {
        long long pxi = cta_x0 + px_local;
        if (pxi >= p.nx) pxi = p.nx - 1;
        px0 = p.xc[3 * pxi];
        px1 = p.xc[3 * pxi + 1];
        px2 = p.xc[3 * pxi + 2];
    }

    double dnu = 0.0, nu0 = 0.0;
    if (!EXACT) {
        if (p.nchan > 1) dnu = (p.freq[p.nchan - 1] - p.freq[0]) / (double)(p.nchan - 1);
        nu0 = p.freq[cta_f0];
    } else {
        for (int i = tid; i < ft; i += NT) fq[i] = p.freq[min(cta_f0 + i, p.nchan - 1)];
    }

    ACC are[CH][NCORR];
    ACC aim[CH][ADJ ? 1 : NCORR];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
#pragma unroll
        for (int c = 0; c < NCORR; ++c) are[j][c] = ACC(0);
#pragma unroll
        for (int c = 0; c < (ADJ ? 1 : NCORR); ++c) aim[j][c] = ACC(0);
    }

    const long long ys = (long long)blockIdx.z * p.ysplit;
    const long long ye = min(p.ny, ys + p.ysplit);
    const int ntiles = (int)((ye - ys + yt - 1) / yt);
    const bool f32dot = p.f32dot != 0;
    const int valid_ch = min(ft, p.nchan - cta_f0);  // channels of this CTA inside the array

    // ---- asynchronous staging: y coordinates of tile t -> ycs[t%3], W(t) -> w_of(t&1)
    // Tile rows are contiguous in shared memory, so granule c of the tile lands at byte
    // c*g; its source is row (c >> rcl) of the tile at byte offset (c & mask)*g.
    unsigned flagbits = 0;  // drop-mask of this thread's granules of the W tile in flight
    const int g = p.granule, rcl = p.row_chunks_log2;
    const int total_chunks = yt << rcl;
    const int valid_bytes = valid_ch * NV * SZ;
    const long long row_pitch = (long long)p.nchan * NV * SZ;  // bytes between W rows
    const char *w_cta = reinterpret_cast<const char *>(p.w) + (long long)cta_f0 * NV * SZ;
    auto issue_yc = [&](int t) {  // 8-byte copies, zero-filled past the end of the slice
        if (tid < yt * 3) {
            const long long gi = (ys + (long long)t * yt) * 3 + tid;
            const bool ok = gi < ye * 3;
            cp_async(smem_addr(ycs + (size_t)(t % 3) * yt * 3 + tid), p.yc + (ok ? gi : 0), 8,
                     ok ? 8 : 0);
        }
    };
    auto issue_w = [&](int t) {
        if (t >= ntiles || !p.fast) return;
        const long long y0 = ys + (long long)t * yt;
        const int rows_valid = (int)min((long long)yt, ye - y0);
        const char *src_tile = w_cta + y0 * row_pitch;
        const unsigned dst_tile = smem_addr(w_of(t & 1));
        unsigned bits = 0;
        int k = 0;
#pragma unroll 2
        for (int c = tid; c < total_chunks; c += NT, ++k) {
            const int yl = c >> rcl;
            const int off = (c - (yl << rcl)) * g;
            const bool ok = (yl < rows_valid) && (off < valid_bytes);
            cp_async(dst_tile + c * g, src_tile + (ok ? yl * row_pitch + off : 0), g, ok ? g : 0);
            if (ADJ && p.anyflag != nullptr && ok) {
                // scalars of this granule -> samples -> drop bits
                const uint8_t *af = p.anyflag + (y0 + yl) * p.nchan + cta_f0;
                const int s0 = off / SZ;
                for (int e = 0; e < g / SZ; ++e)
                    if (af[(s0 + e) / NV]) bits |= 1u << (k * 4 + e);
            }
        }
        flagbits = bits;
    };
    // generic (synchronous) staging: correlation sub-blocks of a wider W
    auto stage_tile_slow = [&](int t) {
        const long long y0 = ys + (long long)t * yt;
        ACC *wt = w_of(t & 1);
        const ACC *wsrc = reinterpret_cast<const ACC *>(p.w);
        const int per_y = ft * NV;
        const int total = yt * per_y;
        for (int idx = tid; idx < total; idx += NT) {
            const int yl = idx / per_y;
            const int rem = idx - yl * per_y;
            const int fl = rem / NV;
            const int e = rem - fl * NV;
            const long long y = y0 + yl;
            const int f = cta_f0 + fl;
            ACC val = ACC(0);
            if (y < ye && f < p.nchan) {
                const long long sample = y * p.nchan + f;
                const bool drop = ADJ && p.anyflag != nullptr && p.anyflag[sample] != 0;
                if (!drop) {
                    const long long base = sample * p.wstride + p.coff;
                    val = WC ? wsrc[2 * (base + (e >> 1)) + (e & 1)] : wsrc[base + e];
                }
            }
            wt[idx] = val;
        }
    };
    auto apply_flags = [&](int t) {
        if (!(ADJ && p.fast) || flagbits == 0) return;
        char *wt = reinterpret_cast<char *>(w_of(t & 1));
        int k = 0;
        for (int c = tid; c < total_chunks; c += NT, ++k) {
            const unsigned bsel = (flagbits >> (k * 4)) & 0xFu;
            if (bsel) {
                ACC *dst = reinterpret_cast<ACC *>(wt + (size_t)c * g);
                for (int e = 0; e < g / SZ; ++e)
                    if (bsel & (1u << e)) dst[e] = ACC(0);
            }
        }
    };

    // ---- anchors of tile `t` into buffer t&1 (y coordinates already in shared memory)
    auto produce_tile = [&](int t) {
        if (t >= ntiles) return;
        const long long y0 = ys + (long long)t * yt;
        const double *yct = ycs + (size_t)(t % 3) * yt * 3;
        CA *anch = anch_of(t & 1);
        CA *dstp = dstp_of(t & 1);
        double *phis = reinterpret_cast<double *>(anch);
        for (int yl = py0; yl < yt; yl += pystep) {
            const long long y = y0 + yl;
            const bool live = y < ye;
            double phi = 0.0;
            if (live)
                phi = __dmul_rn(p.cst, phase_dot(px0, px1, px2, yct[3 * yl], yct[3 * yl + 1],
                                                 yct[3 * yl + 2], f32dot));
            if (EXACT) {
                phis[yl * xgw + px_local] = phi;
            } else {
                C2<double> a = {0.0, 0.0}, d = {0.0, 0.0}, D = {1.0, 0.0}, d8 = {0.0, 0.0};
                if (live) pair_anchors<CH>(phi, nu0, dnu, nck > 1, a, d, D, kSub ? &d8 : nullptr);
                CA dd;
                dd.re = (ACC)d.re;
                dd.im = (ACC)d.im;
                dstp[(yl * xgw + px_local) * DS] = dd;
                if (kSub) {
                    dd.re = (ACC)d8.re;
                    dd.im = (ACC)d8.im;
                    dstp[(yl * xgw + px_local) * DS + 1] = dd;
                }
                store_run_anchors<ACC>(anch + (size_t)yl * nck * xgw + px_local, nck, xgw, a, D);
            }
        }
    };

    auto consume_tile = [&](int t) {
        const CA *anch = anch_of(t & 1);
        const CA *dstp = dstp_of(t & 1);
        const ACC *wt = w_of(t & 1);
        if (EXACT) {
            const double *phis = reinterpret_cast<const double *>(anch);
#pragma unroll 1
            for (int yl = 0; yl < yt; ++yl) {
                const double phi = phis[yl * xgw + x_local];
                const ACC *wrow = wt + (size_t)(yl * ft + fo) * NV;
#pragma unroll
                for (int j = 0; j < CH; j += G) {
                    ACC wv[G * NV];
                    load_vec<G * NV>(wrow + j * NV, wv);
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const C2<double> zd = cis_noinline(__dmul_rn(phi, fq[fo + j + g]));
                        CA z;
                        z.re = (ACC)zd.re;
                        z.im = (ACC)zd.im;
                        accumulate<NCORR, WC, ADJ, ACC>(are[j + g], aim[j + g], z, wv + g * NV);
                    }
                }
            }
        } else {
            // ONE rotation chain per thread: with >= 4 warps per scheduler the 8-cycle DFMA
            // latency is hidden by the other warps, and the single chain lets ptxas pair
            // instructions that share a register operand (.reuse).  A DFMA with three
            // distinct 64-bit register operands issues every 3 cycles on sm_100a (register
            // file bandwidth), one with a reused operand every 2 -- measured, see DESIGN.md.
            // running pointers: every instruction that is not a DFMA still reads the register
            // file, whose bandwidth is what bounds the DFMA stream
            const CA *pa = anch + ck * xgw + x_local;
            const CA *pd = dstp + x_local * DS;
            const ACC *pw = wt + fo * NV;
            const int w_step = ft * NV;
#pragma unroll 1
            for (int yl = 0; yl < yt; ++yl) {
                const CA z = *pa;
                const CA d = pd[0];
                if constexpr (kSub)
                    consume_run<NCORR, WC, ADJ, ACC, CH, G>(are, aim, z, d, pw, pd[1]);
                else
                    consume_run<NCORR, WC, ADJ, ACC, CH, G>(are, aim, z, d, pw);
                pa += NW * 32;  // nck * xgw
                pd += xgw * DS;
                pw += w_step;
            }
        }
    };

    // ---- software pipeline, one barrier per tile
    if (ntiles > 0) {
        issue_yc(0);
        issue_yc(1);
        issue_w(0);
        cp_async_commit();
        if (!p.fast) stage_tile_slow(0);
        cp_async_wait_all();
        apply_flags(0);
        __syncthreads();
        produce_tile(0);
        for (int t = 0; t < ntiles; ++t) {
            cp_async_wait_all();  // this thread's granules of W(t) and y(t+1) have landed
            if (t > 0) apply_flags(t);
            __syncthreads();  // W(t), y(t+1), anchors(t) visible; buffers of tile t-1 free
            issue_yc(t + 2);
            issue_w(t + 1);
            cp_async_commit();
            if (!p.fast && t + 1 < ntiles) stage_tile_slow(t + 1);
            // consume(t) and produce(t+1) are independent: alternate their order between the
            // warps of a scheduler so the latency-bound anchor math of one warp overlaps the
            // FP64-dense rotation loop of its neighbours
            if ((warp >> 2) & 1) {
                produce_tile(t + 1);
                consume_tile(t);
            } else {
                consume_tile(t);
                produce_tile(t + 1);
            }
        }
    }

    // ---- write the owner's channel run
    if (x < p.nx) {
        ACC *o = reinterpret_cast<ACC *>(p.out) + (size_t)blockIdx.z * p.out_split_stride;
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const int f = cta_f0 + fo + j;
            if (f < p.nchan) {
                const long long base = (x * p.nchan + f) * p.wstride + p.coff;
#pragma unroll
                for (int c = 0; c < NCORR; ++c) {
                    if (ADJ) {
                        o[base + c] = are[j][c];
                    } else {
                        C2<ACC> v;
                        v.re = are[j][c];
                        v.im = aim[j][c];
                        reinterpret_cast<C2<ACC> *>(o)[base + c] = v;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Warp-specialised variant: the default for the FP64 kernels (AFR_WS=0 selects the
// single-role kernel above, which the FP32 variants always use).
//
// Same tiling, anchors and rotation loop as phasor_stream_kernel, but the CTA is
// NWC consumer warps + 4 producer warps.  Producers run up to two tiles ahead: they wait
// for a buffer to be released (mbarrier "empty"), cp.async the W tile into it, compute the
// tile's anchors and publish it (mbarrier "full").  A consumer warp waits for "full", runs
// the FP64 rotation loop, and releases the buffer with ONE arrive per warp -- consumer
// warps never synchronise with each other, so a late warp stalls nobody and the integer /
// sincos-heavy producer instructions fill the issue slots the FP64 pipe leaves empty.
// Registers are rebalanced with setmaxnreg inside the CTA's launch-time pool (640 x 96 =
// 512 x 104 + 128 x 64): the extra 8 registers let ptxas keep the operand-reuse-friendly
// schedule of the rotation loop (tools/sass_dp_model.py: 14.2 vs 16.0 cycles per term).
// ---------------------------------------------------------------------------
constexpr int kProducerWarps = 4;

// NCKT: channel runs per CTA as a compile-time constant (0 = runtime).  NCKT == NWC ("full
// width"): the CTA's 16 consumer warps are 16 channel runs of the same 32 owners (nck == NWC,
// the shape of every launch with >= 16 * CH channels): strides become immediates.
// PW producer warps: 4, or 8 for the adjoint with few channels per CTA, where the anchor work per
// (owner, streamed item) pair is spread over so few terms that four producer warps fall behind.
template <int NCORR, bool WC, bool ADJ, typename ACC, int CH, int NWC, bool EXACT, int CREGS, int PREGS,
          int NCKT, int PW = kProducerWarps, bool MMA = false>
__global__ void __launch_bounds__((NWC + PW) * 32, 1)
    phasor_stream_ws_kernel(const DftParams p) {
    static_assert(!MMA || (NCORR == 4 && WC && !ADJ && sizeof(ACC) == 8 && !EXACT && CH == 8),
                  "tensor-pipe consumers: 2x2 complex W, forward, FP64, equispaced channels");
    constexpr int RG = MMA ? 16 / CH : 4;  // tensor-pipe consumers: row groups of 8 owners per warp
    constexpr int NTP = PW * 32;  // producer threads
    constexpr int NV = NCORR * (WC ? 2 : 1);
    constexpr int G = (NV * (int)sizeof(ACC) >= 16) ? 1 : 16 / (NV * (int)sizeof(ACC));
    constexpr int SZ = (int)sizeof(ACC);
    static_assert(CH % G == 0 && (G * NV * SZ) % 16 == 0, "tiling");
    static_assert((CH & (CH - 1)) == 0 && NWC % 4 == 0, "shape");
    using CA = C2<ACC>;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr bool FULLW = NCKT == NWC;
    constexpr bool kUnrollTile = FULLW && !ADJ && NCORR >= 2 && !MMA;  // see the consumer loop
    const int nck = NCKT ? NCKT : p.nck, yt = kUnrollTile ? 8 : p.yt;
    const int xgw = (NWC / nck) * (MMA ? 8 * RG : 32);
    const int ft = nck * CH;
    const int cta_f0 = blockIdx.y * ft;
    const long long cta_x0 = (long long)blockIdx.x * xgw;

    // Tensor-pipe consumers read the anchors / W of TWO streamed items per fragment load: their
    // per-item pitches get 32 bytes of padding, so that the two items fall into different bank
    // groups (unpadded, the pitches are multiples of 128 bytes: two-way conflicts on every load).
    const int apitch = nck * xgw + (MMA ? 2 : 0);  // CA per streamed item
    const int dpitch = xgw + (MMA ? 2 : 0);        // CA per streamed item
    const int wpitch = ft * NV + (MMA ? 4 : 0);    // ACC per streamed item
    const size_t anch_elems = (size_t)yt * apitch;
    const size_t dstp_elems = (size_t)yt * dpitch;
    const size_t w_elems = (size_t)yt * wpitch;
    const size_t buf_bytes = (anch_elems + dstp_elems) * sizeof(CA) + w_elems * SZ;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + 2 * buf_bytes);  // full[2], empty[2], wland[2]
    double *fq = reinterpret_cast<double *>(bars + 6);                        // [ft] (exact)
    auto anch_of = [&](int b) { return reinterpret_cast<CA *>(smem_raw + (size_t)b * buf_bytes); };
    auto dstp_of = [&](int b) { return anch_of(b) + anch_elems; };
    auto w_of = [&](int b) { return reinterpret_cast<ACC *>(dstp_of(b) + dstp_elems); };

    const long long ys = (long long)blockIdx.z * p.ysplit;
    const long long ye = min(p.ny, ys + p.ysplit);
    const int ntiles = (int)((ye - ys + yt - 1) / yt);

    if (tid == 0) {
        mbar_init(&bars[0], NTP);
        mbar_init(&bars[1], NTP);
        // one elected lane per consumer warp releases a buffer (after __syncwarp); with
        // AFR_SANITIZE=1 every lane arrives instead, which compute-sanitizer's racecheck can
        // follow (it does not model the warp-elected release and reports false hazards)
        mbar_init(&bars[2], p.arrive_all ? NWC * 32 : NWC);
        mbar_init(&bars[3], p.arrive_all ? NWC * 32 : NWC);
        // "W tile landed" (flagged adjoint only: producers edit the tile before publishing it)
        mbar_init(&bars[4], 1);
        mbar_init(&bars[5], 1);
    }
    if (EXACT)
        for (int i = tid; i < ft; i += blockDim.x) fq[i] = p.freq[min(cta_f0 + i, p.nchan - 1)];
    __syncthreads();

    if (warp >= NWC) {
        // =============================== PRODUCERS ===============================
        // One producer warp per SM sub-partition shares the issue port with four consumer
        // warps, and its work is a chain of dependent instructions: every instruction it does
        // not execute is latency the consumers do not wait for (ncu: the first version spent
        // 1300 warp-instructions per tile, 80 % of them integer, and the consumers idled 16 %
        // of the time).  Hence: owner coordinates live in registers when a thread always
        // serves the same owner, tile-pair indices are shifts (xgw is a power of two), shared
        // memory addresses advance by a constant, and two pairs are in flight per iteration.
        if (PREGS > 0) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(PREGS > 0 ? PREGS : 24));
        const int ptid = tid - NWC * 32;
        const bool f32dot = p.f32dot != 0;
        const int valid_ch = min(ft, p.nchan - cta_f0);
        const int lxgw = 31 - __clz(xgw);
        const int npairs = yt << lxgw;  // (y, owner) pairs of a tile, y-major
        const bool fixed_owner = xgw <= NTP;  // NTP % xgw == 0: pair q + NTP has the owner of q
        auto load_owner = [&](int xo, double &x0, double &x1, double &x2) {
            long long pxi = cta_x0 + xo;
            if (pxi >= p.nx) pxi = p.nx - 1;
            x0 = p.xc[3 * pxi], x1 = p.xc[3 * pxi + 1], x2 = p.xc[3 * pxi + 2];
        };
        double ox0, ox1, ox2;
        load_owner(ptid & (xgw - 1), ox0, ox1, ox2);
        double dnu = 0.0, nu0 = 0.0;
        if (!EXACT) {
            if (p.nchan > 1) dnu = (p.freq[p.nchan - 1] - p.freq[0]) / (double)(p.nchan - 1);
            nu0 = p.freq[cta_f0];
        }
        const bool need_D = nck > 1;
        const int g = p.granule, rcl = p.row_chunks_log2;
        const int total_chunks = yt << rcl;
        const int valid_bytes = valid_ch * NV * SZ;
        const long long row_pitch = (long long)p.nchan * NV * SZ;
        const char *w_cta = reinterpret_cast<const char *>(p.w) + (long long)cta_f0 * NV * SZ;

        for (int t = 0; t < ntiles; ++t) {
            const int b = t & 1;
            const long long y0 = ys + (long long)t * yt;
            mbar_wait(&bars[2 + b], ((t >> 1) & 1) ^ 1);  // buffer released by every consumer warp
            ACC *wt = w_of(b);
            unsigned long long bits = 0;  // 4 drop bits per granule, <= 16 granules per thread
            uint4 fl16[2];
            if (p.bulk) {
                // W tile by TMA bulk copies, one per row, issued by one thread: the bytes are
                // accounted on the tile's "full" barrier (expect_tx now, this thread's arrival
                // after its anchors), so no producer thread spends instructions on the copy
                const int rows_valid = (int)min((long long)yt, ye - y0);
                const int row_smem = wpitch * SZ;
                const bool edit = ADJ && p.anyflag != nullptr;
                if (ptid == 0) {
                    uint64_t *landed = edit ? &bars[4 + b] : &bars[b];
                    if (edit) {
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(
                                         smem_addr(landed)),
                                     "r"((unsigned)(rows_valid * valid_bytes))
                                     : "memory");
                    } else {
                        mbar_expect_tx(landed, (unsigned)(rows_valid * valid_bytes));
                    }
                    const char *src = w_cta + y0 * row_pitch;
                    char *dst = reinterpret_cast<char *>(wt);
                    for (int yl = 0; yl < rows_valid; ++yl, src += row_pitch, dst += row_smem)
                        bulk_g2s(dst, src, (unsigned)valid_bytes, landed);
                }
                if (edit) {
                    // any-flag bytes of this thread's (<= 2) groups of 16 samples of the tile;
                    // the loads fly while the anchors are computed
                    const int lft = 31 - __clz(ft);
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int s0 = (ptid + i * NTP) * 16;  // first sample of the group
                        const int yl = s0 >> lft, fl = s0 & (ft - 1);
                        fl16[i] = make_uint4(0u, 0u, 0u, 0u);
                        if (yl < rows_valid && fl < valid_ch)
                            fl16[i] = *reinterpret_cast<const uint4 *>(p.anyflag + (y0 + yl) * p.nchan +
                                                                       cta_f0 + fl);
                    }
                }
                // rows past the end of the slice must read as zero (their anchors are zero,
                // but 0 * stale NaN would poison the sum)
                for (int idx = rows_valid * wpitch + ptid; idx < yt * wpitch; idx += NTP) wt[idx] = ACC(0);
            } else if (p.fast) {
                const int rows_valid = (int)min((long long)yt, ye - y0);
                const char *src_tile = w_cta + y0 * row_pitch;
                const unsigned dst_tile = smem_addr(wt);
                int k = 0;
                for (int c = ptid; c < total_chunks; c += NTP, ++k) {
                    const int yl = c >> rcl;
                    const int off = (c - (yl << rcl)) * g;
                    const bool ok = (yl < rows_valid) && (off < valid_bytes);
                    cp_async(dst_tile + c * g, src_tile + (ok ? yl * row_pitch + off : 0), g, ok ? g : 0);
                    if (ADJ && p.anyflag != nullptr && ok) {
                        const uint8_t *af = p.anyflag + (y0 + yl) * p.nchan + cta_f0;
                        const int s0 = off / SZ;
                        for (int e = 0; e < g / SZ; ++e)
                            if (af[(s0 + e) / NV]) bits |= 1ull << (k * 4 + e);
                    }
                }
                cp_async_commit();
            } else {
                const ACC *wsrc = reinterpret_cast<const ACC *>(p.w);
                const int per_y = ft * NV;
                for (int idx = ptid; idx < yt * per_y; idx += NTP) {
                    const int yl = idx / per_y;
                    const int rem = idx - yl * per_y;
                    const int fl = rem / NV;
                    const int e = rem - fl * NV;
                    const long long y = y0 + yl;
                    const int f = cta_f0 + fl;
                    ACC val = ACC(0);
                    if (y < ye && f < p.nchan) {
                        const long long sample = y * p.nchan + f;
                        const bool drop = ADJ && p.anyflag != nullptr && p.anyflag[sample] != 0;
                        if (!drop) {
                            const long long base = sample * p.wstride + p.coff;
                            val = WC ? wsrc[2 * (base + (e >> 1)) + (e & 1)] : wsrc[base + e];
                        }
                    }
                    wt[idx] = val;
                }
            }
            // ---- anchors of tile t: pairs q = ptid, ptid + NTP, ... two per iteration
            CA *anch = anch_of(b);
            CA *dstp = dstp_of(b);
            double *phis = reinterpret_cast<double *>(anch);
            for (int q0 = ptid; q0 < ((MMA && (p.ablate & 1)) ? 0 : npairs); q0 += 2 * NTP) {
                double phi[2];
                bool live[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int q = q0 + i * NTP;
                    const long long y = y0 + (q >> lxgw);
                    live[i] = q < npairs && y < ye;
                    phi[i] = 0.0;
                    if (live[i]) {
                        double x0 = ox0, x1 = ox1, x2 = ox2;
                        if (!fixed_owner) load_owner(q & (xgw - 1), x0, x1, x2);
                        phi[i] = __dmul_rn(p.cst, phase_dot(x0, x1, x2, p.yc[3 * y], p.yc[3 * y + 1],
                                                            p.yc[3 * y + 2], f32dot));
                    }
                }
                if (EXACT) {
                    phis[q0] = phi[0];
                    if (q0 + NTP < npairs) phis[q0 + NTP] = phi[1];
                } else {
                    C2<double> a[2], d[2], D[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        a[i] = {0.0, 0.0}, d[i] = {0.0, 0.0}, D[i] = {1.0, 0.0};
                        if (live[i]) pair_anchors<CH>(phi[i], nu0, dnu, need_D, a[i], d[i], D[i]);
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int q = q0 + i * NTP;
                        if (q < npairs) {
                            CA dd;
                            dd.re = (ACC)d[i].re;
                            dd.im = (ACC)d[i].im;
                            if constexpr (MMA) {
                                dstp[(q >> lxgw) * dpitch + (q & (xgw - 1))] = dd;
                                store_run_anchors<ACC>(anch + (q >> lxgw) * apitch + (q & (xgw - 1)), nck, xgw, a[i], D[i]);
                            } else {
                                dstp[q] = dd;
                                // anchors of pair (yl, xo) start at (yl * nck) * xgw + xo
                                store_run_anchors<ACC>(anch + (((q >> lxgw) * nck) << lxgw) + (q & (xgw - 1)),
                                                       nck, xgw, a[i], D[i]);
                            }
                        }
                    }
                }
            }
            if (ADJ && p.bulk && p.anyflag != nullptr) {
                mbar_wait(&bars[4 + b], (t >> 1) & 1);  // the TMA copies of this tile have landed
                const int lft = 31 - __clz(ft);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if ((fl16[i].x | fl16[i].y | fl16[i].z | fl16[i].w) != 0u) {
                        const unsigned wds[4] = {fl16[i].x, fl16[i].y, fl16[i].z, fl16[i].w};
                        ACC *dst = wt + (size_t)(ptid + i * NTP) * 16 * NV;
#pragma unroll
                        for (int e = 0; e < 16; ++e)
                            if ((wds[e >> 2] >> (8 * (e & 3))) & 0xFFu) {
#pragma unroll
                                for (int v = 0; v < NV; ++v) dst[e * NV + v] = ACC(0);
                            }
                    }
                }
            }
            if (p.fast && !p.bulk) {
                cp_async_wait_all();
                if (ADJ && bits) {  // zero this thread's flagged scalars
                    char *wtb = reinterpret_cast<char *>(wt);
                    int k = 0;
                    for (int c = ptid; c < total_chunks; c += NTP, ++k) {
                        const unsigned bsel = (unsigned)(bits >> (k * 4)) & 0xFu;
                        if (bsel) {
                            ACC *dst = reinterpret_cast<ACC *>(wtb + (size_t)c * g);
                            for (int e = 0; e < g / SZ; ++e)
                                if (bsel & (1u << e)) dst[e] = ACC(0);
                        }
                    }
                }
            }
            mbar_arrive(&bars[b]);  // release: anchors, W tile and flag zeroing are visible
        }
        return;
    }

    // ================================= CONSUMERS =================================
    if (CREGS > 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(CREGS > 0 ? CREGS : 24));
    const int ck = warp % nck;
    const int x_local = (warp / nck) * 32 + lane;
    const long long x = cta_x0 + x_local;
    const int fo = ck * CH;

    if constexpr (MMA) {
        // ---- consumers on the FP64 tensor pipe (2x2 complex W, forward).  For one channel the source
        // sum acc[row][c] += z[row,s] W[s,c] is a complex (rows x sources) x (sources x 4) product; as a
        // real GEMM with the phasor parts as the k index it is exactly the DMMA shape m8n8k4:
        //   A[row][k = 2 s + part] = part ? Im z : Re z                       (8 rows x 2 sources)
        //   B[k][n = 2 c + opart]  = (part ^ opart) ? (+-)Im W[s,c] : Re W[s,c]   (minus for part = 1, opart = 0)
        //   C[row][n] = Re / Im acc[row][c]
        // -- all 256 FMAs of the instruction are useful (16 terms x 16 real FMAs), where the scalar loop
        // spends 16 DFMA per term on a pipe that a DFMA with three distinct register operands only
        // fills to two thirds and 4 broadcast LDS.128 per term (ncu: l1tex 79 %).  Each lane
        // advances ONE real component of z by the three-term recurrence (re and im share the
        // coefficient): at most one DFMA per DMMA, and one double of W per k-step and channel, shared
        // by the warp's RG row groups.
        const int grp = lane >> 2, kq = lane & 3;
        const int s_local = kq >> 1, part = kq & 1;
        const int opart = grp & 1, cb = grp >> 1;
        const int comp = part ^ opart;
        const int bflip = (part & (opart ^ 1)) ? (int)0x80000000 : 0;
        const int xrow0 = (warp / nck) * (8 * RG);
        const int tflip = part ? 0 : (int)0x80000000;
        double acc[RG][CH][2];
#pragma unroll
        for (int g = 0; g < RG; ++g)
#pragma unroll
            for (int j = 0; j < CH; ++j) acc[g][j][0] = acc[g][j][1] = 0.0;
        for (int t = 0; t < ntiles; ++t) {
            const int b = t & 1;
            mbar_wait(&bars[b], (t >> 1) & 1);
            const CA *pa = anch_of(b) + (size_t)s_local * apitch + ck * xgw + xrow0 + grp;
            const CA *pd = dstp_of(b) + (size_t)s_local * dpitch + xrow0 + grp;
            const double *pw = reinterpret_cast<const double *>(w_of(b)) + (size_t)s_local * wpitch + fo * NV + 2 * cb + comp;
            auto bload = [&](int j) {
                const double bw = pw[j * NV];
                return __hiloint2double(__double2hiint(bw) ^ bflip, __double2loint(bw));
            };
            auto mma = [&](double (&c)[2], double a, double bw) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[0]), "+d"(c[1])
                             : "d"(a), "d"(bw));
            };
#pragma unroll 1
            for (int ys = 0; ys < yt; ys += 2) {
                // Scalar FP64 instructions are issued in dense bursts of independent instructions, not
                // sprinkled between the DMMAs: every DMMA -> DFMA -> DMMA switch of the FP64 pipe costs
                // ~8 cycles on top of the 2.2 per DFMA (tools/dmma_mix_microbench.cu: 1 DFMA between
                // DMMAs costs 10 cycles, 16 in a row 2.8 each).
                double xm[RG], xc[RG], c2[RG];  // z_{j-1}, z_j (this lane's part), 2 Re d
                CA z[RG], d[RG];
#pragma unroll
                for (int g = 0; g < RG; ++g) z[g] = pa[8 * g], d[g] = pd[8 * g];
#pragma unroll
                for (int g = 0; g < RG; ++g) xm[g] = z[g].re, xc[g] = z[g].im, c2[g] = d[g].re;  // (ablation only)
                double bw0 = bload(0), bw1 = bload(1);
                // burst 1: z_1 = z_0 d (this lane's part: re = z.re d.re - z.im d.im, im = z.re d.im + z.im d.re)
                // (each burst is the body of a loop that runs once, trip count opaque to the compiler:
                // ptxas schedules inside basic blocks and otherwise spreads the DFMAs between the DMMAs)
#pragma unroll 1
                for (int once = 0; once < ((p.ablate & 2) ? 0 : p.one); ++once) {
                    double tq[RG];
#pragma unroll
                    for (int g = 0; g < RG; ++g) tq[g] = z[g].im * (part ? d[g].re : d[g].im);
#pragma unroll
                    for (int g = 0; g < RG; ++g) c2[g] = d[g].re + d[g].re;
#pragma unroll
                    for (int g = 0; g < RG; ++g) {
                        const double t = __hiloint2double(__double2hiint(tq[g]) ^ tflip, __double2loint(tq[g]));
                        xm[g] = part ? z[g].im : z[g].re;
                        xc[g] = fma(z[g].re, part ? d[g].im : d[g].re, t);
                    }
                }
#pragma unroll
                for (int g = 0; g < RG; ++g) mma(acc[g][0], xm[g], bw0);
#pragma unroll
                for (int g = 0; g < RG; ++g) mma(acc[g][1], xc[g], bw1);
#pragma unroll
                for (int j0 = 2; j0 < CH; j0 += 2) {
                    bw0 = bload(j0), bw1 = bload(j0 + 1);
                    // burst: two recurrence steps of every row group
#pragma unroll 1
                    for (int once = 0; once < ((p.ablate & 2) ? 0 : p.one); ++once) {
#pragma unroll
                        for (int g = 0; g < RG; ++g) xm[g] = fma(c2[g], xc[g], -xm[g]);  // z_{j0}
#pragma unroll
                        for (int g = 0; g < RG; ++g) xc[g] = fma(c2[g], xm[g], -xc[g]);  // z_{j0+1}
                    }
#pragma unroll
                    for (int g = 0; g < RG; ++g) mma(acc[g][j0], xm[g], bw0);
#pragma unroll
                    for (int g = 0; g < RG; ++g) mma(acc[g][j0 + 1], xc[g], bw1);
                }
                pa += 2 * apitch;
                pd += 2 * dpitch;
                pw += 2 * wpitch;
            }
            __syncwarp();
            if (p.arrive_all || lane == 0) mbar_arrive(&bars[2 + b]);
        }
        C2<double> *o = reinterpret_cast<C2<double> *>(reinterpret_cast<double *>(p.out) +
                                                       (size_t)blockIdx.z * p.out_split_stride);
#pragma unroll
        for (int g = 0; g < RG; ++g) {
            const long long xo = cta_x0 + xrow0 + 8 * g + grp;
            if (xo >= p.nx) continue;
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const int f = cta_f0 + fo + j;
                if (f < p.nchan) {
                    C2<double> v;
                    v.re = acc[g][j][0], v.im = acc[g][j][1];
                    o[(xo * p.nchan + f) * p.wstride + p.coff + kq] = v;
                }
            }
        }
        return;
    }

    ACC are[CH][NCORR];
    ACC aim[CH][ADJ ? 1 : NCORR];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
#pragma unroll
        for (int c = 0; c < NCORR; ++c) are[j][c] = ACC(0);
#pragma unroll
        for (int c = 0; c < (ADJ ? 1 : NCORR); ++c) aim[j][c] = ACC(0);
    }

    for (int t = 0; t < ntiles; ++t) {
        const int b = t & 1;
        mbar_wait(&bars[b], (t >> 1) & 1);
        const CA *anch = anch_of(b);
        const CA *dstp = dstp_of(b);
        const ACC *wt = w_of(b);
        if (EXACT) {
            const double *phis = reinterpret_cast<const double *>(anch);
#pragma unroll 1
            for (int yl = 0; yl < yt; ++yl) {
                const double phi = phis[yl * xgw + x_local];
                const ACC *wrow = wt + (size_t)(yl * ft + fo) * NV;
#pragma unroll
                for (int j = 0; j < CH; j += G) {
                    ACC wv[G * NV];
                    load_vec<G * NV>(wrow + j * NV, wv);
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const C2<double> zd = cis_fast(__dmul_rn(phi, fq[fo + j + g]));
                        CA z;
                        z.re = (ACC)zd.re;
                        z.im = (ACC)zd.im;
                        accumulate<NCORR, WC, ADJ, ACC>(are[j + g], aim[j + g], z, wv + g * NV);
                    }
                }
            }
        } else {
            // running pointers: every instruction that is not a DFMA still reads the register
            // file, whose bandwidth is what bounds the DFMA stream
            const CA *pa = anch + ck * xgw + x_local;
            const CA *pd = dstp + x_local;
            const ACC *pw = wt + fo * NV;
            const int w_step = ft * NV;
            // Tile shape known at compile time (8 y items): every offset becomes an immediate.
            // Measured: +4 % for the forward ncorr >= 2 variants; ptxas spills in the others
            // (ncorr = 1: 2.80 vs 2.87 Tterm/s), which keep the rolled loop.
            if constexpr (kUnrollTile) {
#pragma unroll
                for (int yl = 0; yl < 8; ++yl)
                    consume_run<NCORR, WC, ADJ, ACC, CH, G>(are, aim, pa[yl * NWC * 32], pd[yl * 32],
                                                            pw + yl * (NWC * CH * NV));
            } else {
#pragma unroll 1
                for (int yl = 0; yl < yt; ++yl) {
                    const CA z = *pa;
                    const CA d = *pd;
                    consume_run<NCORR, WC, ADJ, ACC, CH, G>(are, aim, z, d, pw);
                    pa += NWC * 32;  // nck * xgw
                    pd += xgw;
                    pw += w_step;
                }
            }
        }
        __syncwarp();
        if (p.arrive_all || lane == 0) mbar_arrive(&bars[2 + b]);  // done with buffer b
    }

    if (x < p.nx) {
        ACC *o = reinterpret_cast<ACC *>(p.out) + (size_t)blockIdx.z * p.out_split_stride;
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const int f = cta_f0 + fo + j;
            if (f < p.nchan) {
                const long long base = (x * p.nchan + f) * p.wstride + p.coff;
#pragma unroll
                for (int c = 0; c < NCORR; ++c) {
                    if (ADJ) {
                        o[base + c] = are[j][c];
                    } else {
                        C2<ACC> v;
                        v.re = are[j][c];
                        v.im = aim[j][c];
                        reinterpret_cast<C2<ACC> *>(o)[base + c] = v;
                    }
                }
            }
        }
    }
}

// out[i] = sum_k partial[k][i] in fixed k order (deterministic), scalars [coff, coff+nc)
// of every (x, f)
template <typename T>
__global__ void reduce_partials_kernel(const T *partial, T *out, long long n_xf, int wstride,
                                       int coff, int nc, int nsplit, long long split_stride) {
    const long long total = n_xf * nc;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long xf = i / nc;
        const int c = (int)(i - xf * nc);
        const long long idx = xf * wstride + coff + c;
        T s = partial[idx];
        for (int k = 1; k < nsplit; ++k) s += partial[(size_t)k * split_stride + idx];
        out[idx] = s;
    }
}

// any-correlation flag per (y, f) sample (dft/kernels.py:136-137)
__global__ void any_flag_kernel(const uint8_t *flags, long long n_samples, int ncorr,
                                uint8_t *any) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_samples;
         i += (long long)gridDim.x * blockDim.x) {
        uint8_t a = 0;
        for (int c = 0; c < ncorr; ++c) a |= flags[i * ncorr + c];
        any[i] = a ? 1 : 0;
    }
}

__global__ void lm_to_lmn_kernel(const double *lm, long long nsrc, int mode, int lm_f32,
                                 double *lmn) {
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s >= nsrc) return;
    const double l = lm[2 * s], m = lm[2 * s + 1];
    double n;
    if (mode == kLmnPhaseClamp) {
        if (lm_f32) {  // rime/phase.py:23-25,42-43 with lm.dtype == float32
            const float lf = (float)l, mf = (float)m;
            float nf = __fsub_rn(__fsub_rn(1.0f, __fmul_rn(lf, lf)), __fmul_rn(mf, mf));
            nf = __fsub_rn(__fsqrt_rn(nf < 0.0f ? 0.0f : nf), 1.0f);
            n = (double)nf;
        } else {
            n = __dsub_rn(__dsub_rn(1.0, __dmul_rn(l, l)), __dmul_rn(m, m));
            n = __dsub_rn(__dsqrt_rn(n < 0.0 ? 0.0 : n), 1.0);
        }
    } else {  // dft/kernels.py:54 : l**2 in lm.dtype, the rest float64, no clamp
        double l2, m2;
        if (lm_f32) {
            const float lf = (float)l, mf = (float)m;
            l2 = (double)__fmul_rn(lf, lf);
            m2 = (double)__fmul_rn(mf, mf);
        } else {
            l2 = __dmul_rn(l, l);
            m2 = __dmul_rn(m, m);
        }
        n = __dsub_rn(__dsqrt_rn(__dsub_rn(__dsub_rn(1.0, l2), m2)), 1.0);
    }
    lmn[3 * s] = l;
    lmn[3 * s + 1] = m;
    lmn[3 * s + 2] = n;
}

// im_to_vis skips pixels that are exactly zero (dft/kernels.py:64), so a source outside the unit
// disc (n = NaN, kernels.py:54 has no clamp) contributes NaN to exactly the (chan, corr) entries
// where its pixel is non-zero and nothing elsewhere -- a zero-padded image grid reaching past
// l^2 + m^2 = 1 is the usual case, and stays finite -- while NaN * 0 would poison every visibility
// here.  Such sources get the coordinates (0, 0, 0) (a finite phasor: their zero pixels add zero)
// and their non-zero pixels are marked in poison[1 + f * ncorr + c] (poison[0] = any), which
// poison_apply_kernel turns into NaN for every row afterwards -- what the reference computes.
// Only threads whose n is not finite read their image row.
template <typename T>
__global__ void mute_nan_sources_kernel(double *lmn, const T *image, long long nsrc, long long nfc,
                                        int image_complex, uint8_t *poison) {
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s >= nsrc) return;
    const double q = lmn[3 * s] + lmn[3 * s + 1] + lmn[3 * s + 2];
    if (q - q == 0.0) return;  // finite
    const T *row = image + s * nfc * (image_complex ? 2 : 1);
    for (long long i = 0; i < nfc; ++i) {
        const bool lit = image_complex ? (row[2 * i] != T(0) || row[2 * i + 1] != T(0)) : row[i] != T(0);
        if (lit) poison[1 + i] = 1, poison[0] = 1;
    }
    lmn[3 * s] = lmn[3 * s + 1] = lmn[3 * s + 2] = 0.0;
}

template <typename T>
__global__ void poison_apply_kernel(T *out, const uint8_t *poison, long long nrow, long long nfc) {
    if (!poison[0]) return;
    const T nan = T(NAN);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nrow * nfc;
         i += (long long)gridDim.x * blockDim.x)
        if (poison[1 + i % nfc]) out[2 * i] = nan, out[2 * i + 1] = nan;
}

int ilog2(long long v) {
    int r = 0;
    while ((1LL << (r + 1)) <= v) ++r;
    return r;
}

// MMA: consumers on the FP64 tensor pipe (2x2 complex W, forward, FP64, equispaced channels; the
// caller checked that the W rows are 16-byte aligned and cover every correlation)
template <int NCORR, bool WC, bool ADJ, typename ACC, int CH, int NW, bool MMA = false>
int launch_one(DftParams p, bool exact, cudaStream_t stream) {
    constexpr int NV = NCORR * (WC ? 2 : 1);
    constexpr int SZ = (int)sizeof(ACC);
    constexpr int NT = NW * 32;
    const char *ws_env = getenv("AFR_WS");
    // Which FP64 variants run warp-specialised (16 consumer + 4 producer warps) was decided
    // by measurement (tools/ws_sweep.py, B200): the forward (complex-accumulator) kernels
    // gain 5-10 %, the adjoint kernels (64 real accumulators per thread) lose.  AFR_WS=0 / 1
    // forces the choice.
    // The adjoint kernels win too once the W tile arrives by TMA (ADJ c=1: 2.47 vs 2.01
    // Tterm/s), which needs 16-byte rows and, with flags, 16-sample groups.
    constexpr int CHV = (NV * SZ >= 16) ? 1 : 16 / (NV * SZ);
    const bool rows16 = ((long long)p.nchan * NV * SZ) % 16 == 0 &&
                        reinterpret_cast<uintptr_t>(p.w) % 16 == 0 && p.wstride == NCORR && p.coff == 0;
    int nck_ws = 1;
    while (nck_ws < (p.nchan + CH - 1) / CH && nck_ws < NW) nck_ws *= 2;
    const bool flags16 = p.anyflag == nullptr || (p.nchan % 16 == 0 && (nck_ws * CH) % 16 == 0);
    const bool bulk_ok = rows16 && flags16;
    (void)CHV;
    // Anchor work per tile is fixed per (x,y) pair, consumer work grows with the channels a CTA
    // covers and the FP64 instructions per term.  When their ratio is small the 4 producer
    // warps fall behind the 16 consumers and the single-role kernel (all 16 warps share the
    // anchor work) wins.  Measured crossover, c=1: 128 channels per CTA (forward 2.27 vs 2.24,
    // adjoint 2.21 vs 1.89 Tterm/s) -> warp-specialised; 64 channels (1.61 vs 2.02, 1.55 vs
    // 1.83) -> single-role; c=4 forward at 64 channels per CTA -> warp-specialised (0.90 vs 0.73).
    constexpr int kDpPerTerm = NCORR * (WC ? 2 : 1) * (ADJ ? 1 : 2) + 2;
    const bool kPreferWS = (!ADJ || bulk_ok) && nck_ws * CH * kDpPerTerm >= 500;
    // FP32 variants gain nothing from it (measured 3.10 vs 3.07 Tterm/s): they are bound by
    // the register-file bandwidth of three-operand FFMAs, not by the anchor work
    // Below the crossover the ADJOINT still gains from dedicated producers if there are eight producer
    // warps (24 warps: 512 x 88 + 256 x 64 registers): configs[4] has 64 channels.  AFR_WS8=0 disables.
    const char *ws8_env = getenv("AFR_WS8");
    const bool use_ws8 = ADJ && sizeof(ACC) == 8 && bulk_ok && !exact && !ws_env && !kPreferWS && nck_ws <= NW / 2 &&
                         nck_ws * CH >= 32 && !(ws8_env && atoi(ws8_env) == 0);
    const bool use_ws = (sizeof(ACC) == 8) && ((ws_env ? atoi(ws_env) != 0 : kPreferWS) || use_ws8 || MMA);
    // channel runs per CTA: the single-role kernel keeps nck <= NW/2 so that every thread
    // owns an (x,y) pair per tile; dedicated producers do not need that, and anchors are
    // cheaper per term the more channels a CTA covers (measured +9 % at 16 runs vs 8)
    const int runs = (p.nchan + CH - 1) / CH;
    int nck_max = use_ws ? NW : NW / 2;
    if (MMA && getenv("AFR_POINT_MMA_NCK")) nck_max = std::max(1, std::min(NW, atoi(getenv("AFR_POINT_MMA_NCK"))));
    int nck = 1;
    while (nck < runs && nck < nck_max) nck *= 2;
    p.nck = nck;
    p.arrive_all = (getenv("AFR_SANITIZE") && atoi(getenv("AFR_SANITIZE")) != 0) ? 1 : 0;
    p.one = 1;
    p.ablate = getenv("AFR_POINT_MMA_ABLATE") ? atoi(getenv("AFR_POINT_MMA_ABLATE")) : 0;
    const int xgw = (NW / nck) * (MMA ? 8 * (16 / CH) : 32);
    const int ft = nck * CH;
    const long long gx = (p.nx + xgw - 1) / xgw;
    const int gy = (p.nchan + ft - 1) / ft;
    AFR_REQUIRE(gx <= 2147483647LL && gy <= 65535, "phasor_stream: grid too large");

    // W staging mode: cp.async granules when this launch covers every correlation of W
    p.fast = (p.wstride == NCORR && p.coff == 0) ? 1 : 0;
    const long long row_bytes_global = (long long)p.nchan * NV * SZ;
    int granule = 16;
    while (granule > SZ && ((row_bytes_global % granule) != 0 ||
                            (reinterpret_cast<uintptr_t>(p.w) % granule) != 0))
        granule /= 2;
    p.granule = granule;
    const int row_chunks = ft * NV * SZ / granule;
    p.row_chunks_log2 = ilog2(row_chunks);

    // y items per tile: as many (even, <= 8) as fit double-buffered in shared memory and in
    // kMaxChunks cp.async granules per thread
    const size_t per_y = (size_t)(nck + 1 + (kSubRuns<ACC, CH> ? 1 : 0)) * xgw * sizeof(C2<ACC>) +
                         (size_t)ft * NV * SZ + (MMA ? 96 : 0);  // MMA: 32 bytes of padding per array
    // granules one thread may have in flight: 8 per thread of the whole CTA, or 16 per
    // producer thread of the warp-specialised kernel
    const long long max_chunks = use_ws ? 16LL * (use_ws8 ? 8 : kProducerWarps) * 32 : (long long)kMaxChunks * NT;
    // (the granule limit does not apply when the tile will arrive by TMA bulk copies)
    const bool will_bulk = use_ws && p.fast && granule == 16 && bulk_ok;
    int yt = 8;
    while (yt > 2 && (2 * yt * per_y > 200 * 1024 || (!will_bulk && (long long)yt * row_chunks > max_chunks)))
        yt -= 2;
    if (!will_bulk && (long long)yt * row_chunks > max_chunks) p.fast = 0;
    p.yt = yt;
    // TMA bulk copies need 16-byte aligned rows and no flag post-processing of the tile
    p.bulk = (use_ws && p.fast && granule == 16 && bulk_ok) ? 1 : 0;
    if constexpr (MMA) AFR_REQUIRE(p.bulk, "phasor_stream (tensor-pipe consumers): W tile must arrive by bulk copies");
    const size_t smem = 2 * yt * per_y + 3 * (size_t)yt * 3 * sizeof(double) + (size_t)ft * sizeof(double);

    // split the streamed axis when the owners alone cannot fill the machine; pick the
    // split count whose CTA total wastes the least of the last wave (1 CTA per SM)
    const int sms = sm_count();
    long long nsplit = 1;
    const long long ctas = gx * gy;
    const long long max_split = std::max(1LL, (p.ny + 4LL * yt - 1) / (4LL * yt));
    if (ctas < 6LL * sms && max_split > 1) {
        const long long lo = std::max(1LL, (3LL * sms + ctas - 1) / ctas);
        double best = -1.0;
        for (long long ns = lo; ns <= std::min(max_split, lo + 24); ++ns) {
            const long long tot = ctas * ns;
            const long long waves = (tot + sms - 1) / sms;
            const double eff = (double)tot / (double)(waves * sms);
            if (eff > best + 1e-9) {
                best = eff;
                nsplit = ns;
            }
        }
        if (nsplit > max_split) nsplit = max_split;
    }
    long long ysplit = (p.ny + nsplit - 1) / nsplit;
    ysplit = ((ysplit + yt - 1) / yt) * yt;
    nsplit = (p.ny + ysplit - 1) / ysplit;
    AFR_REQUIRE(nsplit <= 65535, "phasor_stream: too many y slices");
    p.ysplit = ysplit;

    const size_t out_scalars = (size_t)p.nx * p.nchan * p.wstride * (ADJ ? 1 : 2);
    Scratch partial;
    void *final_out = p.out;
    if (nsplit > 1) {
        AFR_CUDA_OK(partial.alloc(out_scalars * sizeof(ACC) * nsplit, stream));
        p.out = partial.ptr;
        p.out_split_stride = (long long)out_scalars;
    } else {
        p.out_split_stride = 0;
    }

    dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)nsplit);
    note_dft_path((use_ws ? 1 : 0) | (exact ? 2 : 0) | (p.bulk ? 4 : 0) | (sizeof(ACC) == 4 ? 8 : 0) |
                  (use_ws8 ? 16 : 0) | (MMA ? 32 : 0) | (nck << 8) |
                  ((int)nsplit << 16));
    if constexpr (sizeof(ACC) == 8) {
      if (use_ws) {
        // setmaxnreg can only move registers inside the CTA's launch-time pool: 640 threads
        // x 96 registers = 512 x 104 (consumers) + 128 x 64 (producers)
        constexpr int CREGS = 104, PREGS = 64;
        const size_t smem_ws = 2 * yt * per_y + 6 * sizeof(uint64_t) + (size_t)ft * sizeof(double);
        const int threads = (NW + kProducerWarps) * 32;
        auto launch_ws = [&](auto kern) -> int {
            cudaFuncAttributes attr;
            AFR_CUDA_OK(cudaFuncGetAttributes(&attr, kern));
            // the register move must balance exactly, otherwise setmaxnreg.inc never returns
            AFR_REQUIRE(threads * attr.numRegs >= NW * 32 * CREGS + kProducerWarps * 32 * PREGS,
                        "phasor_stream_ws: launch-time register pool too small for setmaxnreg");
            AFR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem_ws));
            kern<<<grid, threads, smem_ws, stream>>>(p);
            return 0;
        };
        int rc;
        bool launched = false;
        rc = 0;
        if constexpr (ADJ) {
            if (use_ws8) {
                constexpr int PW8 = 8, C8 = 88, P8 = 64;
                auto kern = phasor_stream_ws_kernel<NCORR, WC, ADJ, ACC, CH, NW, false, C8, P8, 0, PW8>;
                const int threads8 = (NW + PW8) * 32;
                cudaFuncAttributes attr;
                AFR_CUDA_OK(cudaFuncGetAttributes(&attr, kern));
                AFR_REQUIRE(threads8 * attr.numRegs >= NW * 32 * C8 + PW8 * 32 * P8,
                            "phasor_stream_ws (8 producer warps): launch-time register pool too small");
                AFR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ws));
                kern<<<grid, threads8, smem_ws, stream>>>(p);
                launched = true;
            }
        }
        if (launched) {
        } else if (exact) {
            rc = launch_ws(phasor_stream_ws_kernel<NCORR, WC, ADJ, ACC, CH, NW, true, CREGS, PREGS, 0>);
            launched = true;
        } else if (MMA) {
            if constexpr (MMA) {
                if (nck == NW)
                    rc = launch_ws(phasor_stream_ws_kernel<NCORR, WC, ADJ, ACC, CH, NW, false, CREGS, PREGS, NW,
                                                           kProducerWarps, true>);
                else if (nck == NW / 2)
                    rc = launch_ws(phasor_stream_ws_kernel<NCORR, WC, ADJ, ACC, CH, NW, false, CREGS, PREGS, NW / 2,
                                                           kProducerWarps, true>);
                else
                    rc = launch_ws(phasor_stream_ws_kernel<NCORR, WC, ADJ, ACC, CH, NW, false, CREGS, PREGS, 0,
                                                           kProducerWarps, true>);
            }
            launched = true;
        } else if (nck == NW && yt == 8) {
            rc = launch_ws(phasor_stream_ws_kernel<NCORR, WC, ADJ, ACC, CH, NW, false, CREGS, PREGS, NW>);
            launched = true;
        } else if (nck == NW / 2) {
            // the adjoint variants have long runs (CH = 32 / 16 / 8): 256 channels are 8 runs
            if constexpr (ADJ) {
                rc = launch_ws(phasor_stream_ws_kernel<NCORR, WC, ADJ, ACC, CH, NW, false, CREGS, PREGS, NW / 2>);
                launched = true;
            }
        }
        if (!launched)
            rc = launch_ws(phasor_stream_ws_kernel<NCORR, WC, ADJ, ACC, CH, NW, false, CREGS, PREGS, 0>);
        if (rc) return rc;
      }
    }
    if (use_ws) {
        // launched above
    } else if (exact) {
        auto kern = phasor_stream_kernel<NCORR, WC, ADJ, ACC, CH, NW, true>;
        AFR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
        kern<<<grid, NT, smem, stream>>>(p);
    } else {
        auto kern = phasor_stream_kernel<NCORR, WC, ADJ, ACC, CH, NW, false>;
        AFR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
        kern<<<grid, NT, smem, stream>>>(p);
    }
    AFR_LAUNCH_OK();

    if (nsplit > 1) {
        const long long n_xf = p.nx * p.nchan;
        const int sc = ADJ ? 1 : 2;  // scalars per correlation
        const long long total = n_xf * NCORR * sc;
        int blocks = (int)std::min<long long>((total + 255) / 256, 8LL * sms);
        if (blocks < 1) blocks = 1;
        reduce_partials_kernel<ACC><<<blocks, 256, 0, stream>>>(
            reinterpret_cast<const ACC *>(partial.ptr), reinterpret_cast<ACC *>(final_out), n_xf,
            p.wstride * sc, p.coff * sc, NCORR * sc, (int)nsplit, (long long)out_scalars);
        AFR_LAUNCH_OK();
    }
    return 0;
}

// Tile shapes per variant: (CH channels per run, NW warps per CTA) chosen so the
// accumulators fit the register budget of NW*32 threads on one SM.
template <bool WC, bool ADJ, typename ACC>
int launch_corr_blocks(DftParams p, bool exact, cudaStream_t stream) {
    constexpr bool F32 = sizeof(ACC) == 4;
    // correlations are handled in blocks of 4, 2, 1 (ncorr = 3 -> 2 + 1, etc.)
    int c = 0;
    const int ncorr = p.wstride;
    while (c < ncorr) {
        p.coff = c;
        int rc;
        if (ncorr - c >= 4) {
            constexpr int CH = ADJ ? (F32 ? 16 : 8) : (F32 ? 8 : 4);
            // 2x2 complex W, forward, FP64, equispaced channels, 16-byte W rows: tensor-pipe consumers
            // (AFR_POINT_MMA=0 keeps the scalar loop)
            const char *mma_env = getenv("AFR_POINT_MMA"), *ws_env = getenv("AFR_WS");
            const bool mma = WC && !ADJ && !F32 && !exact && ncorr == 4 &&
                             reinterpret_cast<uintptr_t>(p.w) % 16 == 0 && !(mma_env && atoi(mma_env) == 0) &&
                             !(ws_env && atoi(ws_env) == 0);
            if constexpr (WC && !ADJ && !F32) {
                // runs of 8 channels x 2 row groups per warp: measured against 4 x 4 (more anchors per
                // term, 511-636 Gterms/s) and 16 x 1 (8-row tiles re-read W too often, 414): 744
                if (mma)
                    rc = launch_one<4, WC, ADJ, ACC, 8, 16, true>(p, exact, stream);
                else
                    rc = launch_one<4, WC, ADJ, ACC, CH, 16>(p, exact, stream);
            } else {
                rc = launch_one<4, WC, ADJ, ACC, CH, 16>(p, exact, stream);
            }
            c += 4;
        } else if (ncorr - c >= 2) {
            constexpr int CH = ADJ ? (F32 ? 32 : 16) : (F32 ? 16 : 8);
            rc = launch_one<2, WC, ADJ, ACC, CH, 16>(p, exact, stream);
            c += 2;
        } else {
            constexpr int CH = ADJ ? 32 : (F32 ? 32 : 16);
            rc = launch_one<1, WC, ADJ, ACC, CH, 16>(p, exact, stream);
            c += 1;
        }
        if (rc) return rc;
    }
    return 0;
}

}  // namespace

int launch_lm_to_lmn(const double *lm, int64_t nsrc, int mode, bool lm_f32, double *lmn,
                     cudaStream_t stream) {
    if (nsrc <= 0) return 0;
    const int blocks = (int)((nsrc + 255) / 256);
    lm_to_lmn_kernel<<<blocks, 256, 0, stream>>>(lm, nsrc, mode, lm_f32 ? 1 : 0, lmn);
    AFR_LAUNCH_OK();
    return 0;
}

int run_phasor_stream(const double *xc, int64_t nx, const double *yc, int64_t ny, const void *w,
                      bool w_complex, const uint8_t *flags, const double *freq, int64_t nchan,
                      int64_t ncorr, double cst, bool f32dot, bool adjoint, bool exact,
                      bool acc32, void *out, cudaStream_t stream) {
    AFR_REQUIRE(nchan <= 2147483647LL / 64 && ncorr <= 64, "phasor_stream: nchan/ncorr too large");
    const size_t out_bytes = (size_t)nx * nchan * ncorr * (adjoint ? 1 : 2) * (acc32 ? 4 : 8);
    if (nx <= 0 || nchan <= 0 || ncorr <= 0) return 0;
    if (ny <= 0) {  // empty sum: the reference returns zeros
        AFR_CUDA_OK(cudaMemsetAsync(out, 0, out_bytes, stream));
        return 0;
    }
    Scratch anyflag;
    if (flags != nullptr) {
        const long long n = ny * nchan;
        AFR_CUDA_OK(anyflag.alloc((size_t)n, stream));
        const int blocks = (int)std::min<long long>((n + 255) / 256, 16LL * sm_count());
        any_flag_kernel<<<blocks, 256, 0, stream>>>(flags, n, (int)ncorr, (uint8_t *)anyflag.ptr);
        AFR_LAUNCH_OK();
    }
    DftParams p{};
    p.xc = xc;
    p.yc = yc;
    p.w = w;
    p.anyflag = (const uint8_t *)anyflag.ptr;
    p.freq = freq;
    p.out = out;
    p.cst = cst;
    p.nx = nx;
    p.ny = ny;
    p.nchan = (int)nchan;
    p.wstride = (int)ncorr;
    p.f32dot = f32dot ? 1 : 0;
    if (!acc32) {
        if (w_complex)
            return adjoint ? launch_corr_blocks<true, true, double>(p, exact, stream)
                           : launch_corr_blocks<true, false, double>(p, exact, stream);
        return adjoint ? launch_corr_blocks<false, true, double>(p, exact, stream)
                       : launch_corr_blocks<false, false, double>(p, exact, stream);
    }
    if (w_complex)
        return adjoint ? launch_corr_blocks<true, true, float>(p, exact, stream)
                       : launch_corr_blocks<true, false, float>(p, exact, stream);
    return adjoint ? launch_corr_blocks<false, true, float>(p, exact, stream)
                   : launch_corr_blocks<false, false, float>(p, exact, stream);
}

}  // namespace afr

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
using namespace afr;

extern "C" int afr_im_to_vis(const void *image, int image_complex, const double *uvw,
                             const double *lm, const double *freq, int64_t nsrc, int64_t nrow,
                             int64_t nchan, int64_t ncorr, int convention, int f32_flags,
                             int chan_mode, int out_c64, void *out, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AFR_REQUIRE(convention == AFR_FOURIER || convention == AFR_CASA,
                "convention not in ('fourier', 'casa')");
    AFR_REQUIRE(nsrc >= 0 && nrow >= 0 && nchan >= 0 && ncorr >= 0, "negative extent");
    // dft/kernels.py:34-39
    const double cst = convention == AFR_FOURIER ? -kTwoPiOverC : kTwoPiOverC;
    Scratch lmn;
    AFR_CUDA_OK(lmn.alloc(sizeof(double) * 3 * (size_t)nsrc, stream));
    int rc = launch_lm_to_lmn(lm, nsrc, kLmnDft, (f32_flags & AFR_F32_LM) != 0,
                              (double *)lmn.ptr, stream);
    if (rc) return rc;
    const long long nfc = nchan * ncorr;
    Scratch poison;
    if (nsrc > 0 && nfc > 0) {
        AFR_CUDA_OK(poison.alloc((size_t)nfc + 1, stream));
        AFR_CUDA_OK(cudaMemsetAsync(poison.ptr, 0, (size_t)nfc + 1, stream));
        const unsigned blocks = (unsigned)((nsrc + 255) / 256);
        if (out_c64)
            mute_nan_sources_kernel<float><<<blocks, 256, 0, stream>>>(
                (double *)lmn.ptr, (const float *)image, nsrc, nfc, image_complex, (uint8_t *)poison.ptr);
        else
            mute_nan_sources_kernel<double><<<blocks, 256, 0, stream>>>(
                (double *)lmn.ptr, (const double *)image, nsrc, nfc, image_complex, (uint8_t *)poison.ptr);
        AFR_LAUNCH_OK();
    }
    const bool f32dot = (f32_flags & AFR_F32_LM) && (f32_flags & AFR_F32_UVW);
    rc = run_phasor_stream(uvw, nrow, (const double *)lmn.ptr, nsrc, image, image_complex != 0,
                           nullptr, freq, nchan, ncorr, cst, f32dot, /*adjoint=*/false,
                           chan_mode == AFR_CHAN_EXACT, out_c64 != 0, out, stream);
    if (rc) return rc;
    if (nsrc > 0 && nfc > 0 && nrow > 0) {
        const unsigned blocks = (unsigned)std::min<long long>((nrow * nfc + 255) / 256, 4LL * sm_count());
        if (out_c64)
            poison_apply_kernel<float><<<blocks, 256, 0, stream>>>((float *)out, (const uint8_t *)poison.ptr,
                                                                   nrow, nfc);
        else
            poison_apply_kernel<double><<<blocks, 256, 0, stream>>>((double *)out, (const uint8_t *)poison.ptr,
                                                                    nrow, nfc);
        AFR_LAUNCH_OK();
    }
    return 0;
}

extern "C" int afr_vis_to_im(const void *vis, int vis_complex, const double *uvw,
                             const double *lm, const double *freq, const uint8_t *flags,
                             int64_t nsrc, int64_t nrow, int64_t nchan, int64_t ncorr,
                             int convention, int f32_flags, int chan_mode, int out_f32, void *out,
                             void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AFR_REQUIRE(convention == AFR_FOURIER || convention == AFR_CASA,
                "convention not in ('fourier', 'casa')");
    AFR_REQUIRE(nsrc >= 0 && nrow >= 0 && nchan >= 0 && ncorr >= 0, "negative extent");
    // dft/kernels.py:110-115 : opposite sign to im_to_vis
    const double cst = convention == AFR_FOURIER ? kTwoPiOverC : -kTwoPiOverC;
    Scratch lmn;
    AFR_CUDA_OK(lmn.alloc(sizeof(double) * 3 * (size_t)nsrc, stream));
    int rc = launch_lm_to_lmn(lm, nsrc, kLmnDft, (f32_flags & AFR_F32_LM) != 0,
                              (double *)lmn.ptr, stream);
    if (rc) return rc;
    const bool f32dot = (f32_flags & AFR_F32_LM) && (f32_flags & AFR_F32_UVW);
    return run_phasor_stream((const double *)lmn.ptr, nsrc, uvw, nrow, vis, vis_complex != 0,
                             flags, freq, nchan, ncorr, cst, f32dot, /*adjoint=*/true,
                             chan_mode == AFR_CHAN_EXACT, out_f32 != 0, out, stream);
}
