// predict_vis / apply_gains (africanus/rime/predict.py:466-649) and the fused
// phase_delay (x) brightness -> predict_vis composition
// (africanus/rime/examples/predict.py:107-134,490,522-527) for sm_100a.
//
//   V[r,f] = G1[t,a1,f] (B[r,f] + sum_s E1[s,t,a1,f] X[s,r,f] E2[s,t,a2,f]^H) G2[t,a2,f]^H
//
// * predict_vis with a materialised source_coh is bound by the HBM read of X
//   (16*ncorr B/term): one thread owns one (row, chan) output, lanes run along
//   chan so every load/store is contiguous across the warp, the 2x2 chain stays in
//   registers, the source loop is sequential (the reference's accumulation order).
// * the fused kernels never materialise K or X:
//     - without DDEs the source sum is the phasor-stream kernel of afr_dft.cu with the
//       brightness as a complex "image" (CH-channel runs per thread, anchored rotation);
//     - with DDEs a warp owns (row, 32*kK channels): lane L handles channels
//       f0+L, f0+L+32, ... so DDE gathers stay coalesced along chan, and the phasor
//       advances by exp(i*phi*32*dnu) between a lane's channels.
//   base_vis and the DIE product are applied by the same epilogue as predict_vis.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "afr_dft.cuh"

namespace afr {
namespace {

template <typename T>
struct Cx {
    T re, im;
};

template <typename T>
__device__ __forceinline__ Cx<T> mul(Cx<T> a, Cx<T> b) {
    return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
template <typename T>
__device__ __forceinline__ Cx<T> mulc(Cx<T> a, Cx<T> b) {  // a * conj(b)
    return {a.re * b.re + a.im * b.im, a.im * b.re - a.re * b.im};
}
template <typename T>
__device__ __forceinline__ Cx<T> add(Cx<T> a, Cx<T> b) {
    return {a.re + b.re, a.im + b.im};
}

template <typename T, int N>
__device__ __forceinline__ void load_n(const Cx<T> *p, Cx<T> (&v)[N]) {
    // 16-byte vector loads where the element size allows
    if (sizeof(Cx<T>) == 16) {
        const double2 *q = reinterpret_cast<const double2 *>(p);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double2 d = q[i];
            v[i].re = (T)d.x;
            v[i].im = (T)d.y;
        }
    } else if (N % 2 == 0) {
        const float4 *q = reinterpret_cast<const float4 *>(p);
#pragma unroll
        for (int i = 0; i < N / 2; ++i) {
            float4 d = q[i];
            v[2 * i].re = (T)d.x;
            v[2 * i].im = (T)d.y;
            v[2 * i + 1].re = (T)d.z;
            v[2 * i + 1].im = (T)d.w;
        }
    } else {
        const float2 *q = reinterpret_cast<const float2 *>(p);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            float2 d = q[i];
            v[i].re = (T)d.x;
            v[i].im = (T)d.y;
        }
    }
}

template <typename T, int N>
__device__ __forceinline__ void store_n(Cx<T> *p, const Cx<T> (&v)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = v[i];
}

// out (+)= A * (X * C^H), association of predict.py:103-117
template <typename T>
__device__ __forceinline__ void jones3_2x2(const Cx<T> (&a)[4], const Cx<T> (&x)[4],
                                           const Cx<T> (&c)[4], Cx<T> (&o)[4], bool accumulate) {
    const Cx<T> xx = add(mulc(x[0], c[0]), mulc(x[1], c[1]));
    const Cx<T> xy = add(mulc(x[0], c[2]), mulc(x[1], c[3]));
    const Cx<T> yx = add(mulc(x[2], c[0]), mulc(x[3], c[1]));
    const Cx<T> yy = add(mulc(x[2], c[2]), mulc(x[3], c[3]));
    const Cx<T> r0 = add(mul(a[0], xx), mul(a[1], yx));
    const Cx<T> r1 = add(mul(a[0], xy), mul(a[1], yy));
    const Cx<T> r2 = add(mul(a[2], xx), mul(a[3], yx));
    const Cx<T> r3 = add(mul(a[2], xy), mul(a[3], yy));
    if (accumulate) {
        o[0] = add(o[0], r0);
        o[1] = add(o[1], r1);
        o[2] = add(o[2], r2);
        o[3] = add(o[3], r3);
    } else {
        o[0] = r0;
        o[1] = r1;
        o[2] = r2;
        o[3] = r3;
    }
}

// out += A * C^H  (predict.py:138-148)
template <typename T>
__device__ __forceinline__ void jones2_2x2(const Cx<T> (&a)[4], const Cx<T> (&c)[4],
                                           Cx<T> (&o)[4]) {
    o[0] = add(o[0], add(mulc(a[0], c[0]), mulc(a[1], c[1])));
    o[1] = add(o[1], add(mulc(a[0], c[2]), mulc(a[1], c[3])));
    o[2] = add(o[2], add(mulc(a[2], c[0]), mulc(a[3], c[1])));
    o[3] = add(o[3], add(mulc(a[2], c[2]), mulc(a[3], c[3])));
}

struct PredictParams {
    const int32_t *time_index, *ant1, *ant2;
    const void *dde1, *coh, *dde2, *die1, *bvis, *die2;
    const void *acc_init;  // optional (row,chan,C) pre-summed coherencies (fused path)
    void *out;
    long long nsrc, nrow, ntime, nant;
    long long nfc;  // chan (2x2 mode) or chan*ncorr (diagonal mode, element-wise)
};

// N = 4: (2,2) matrices; N = 1: element-wise ("diagonal" (1,) / (2,) Jones, flattened
// over chan*corr because every correlation is independent, predict.py:93-98)
template <typename T, int N>
__global__ void __launch_bounds__(256) predict_vis_kernel(const PredictParams p) {
    using C = Cx<T>;
    const C *dde1 = (const C *)p.dde1, *dde2 = (const C *)p.dde2, *coh = (const C *)p.coh;
    const C *die1 = (const C *)p.die1, *die2 = (const C *)p.die2, *bvis = (const C *)p.bvis;
    const C *init = (const C *)p.acc_init;
    C *out = (C *)p.out;
    const bool have_dde = dde1 != nullptr, have_coh = coh != nullptr;
    const long long total = p.nrow * p.nfc;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / p.nfc, f = i - r * p.nfc;
        const long long ti = p.time_index[r], a1 = p.ant1[r], a2 = p.ant2[r];
        C acc[N];
#pragma unroll
        for (int k = 0; k < N; ++k) acc[k] = {T(0), T(0)};
        if (init) load_n<T, N>(init + i * N, acc);
        if (have_dde || have_coh) {
            for (long long s = 0; s < p.nsrc; ++s) {
                C x[N], e1[N], e2[N];
                if (have_coh) load_n<T, N>(coh + ((s * p.nrow + r) * p.nfc + f) * N, x);
                if (have_dde) {
                    load_n<T, N>(dde1 + (((s * p.ntime + ti) * p.nant + a1) * p.nfc + f) * N, e1);
                    load_n<T, N>(dde2 + (((s * p.ntime + ti) * p.nant + a2) * p.nfc + f) * N, e2);
                }
                if constexpr (N == 4) {
                    if (have_dde && have_coh) {
                        jones3_2x2<T>((const C(&)[4])e1, (const C(&)[4])x, (const C(&)[4])e2,
                                      (C(&)[4])acc, true);
                    } else if (have_dde) {
                        jones2_2x2<T>((const C(&)[4])e1, (const C(&)[4])e2, (C(&)[4])acc);
                    } else {
#pragma unroll
                        for (int k = 0; k < N; ++k) acc[k] = add(acc[k], x[k]);
                    }
                } else {
                    if (have_dde && have_coh)
                        acc[0] = add(acc[0], mulc(mul(e1[0], x[0]), e2[0]));
                    else if (have_dde)
                        acc[0] = add(acc[0], mulc(e1[0], e2[0]));
                    else
                        acc[0] = add(acc[0], x[0]);
                }
            }
        }
        if (bvis) {
            C b[N];
            load_n<T, N>(bvis + i * N, b);
#pragma unroll
            for (int k = 0; k < N; ++k) acc[k] = add(acc[k], b[k]);
        }
        if (die1) {
            C g1[N], g2[N];
            load_n<T, N>(die1 + ((ti * p.nant + a1) * p.nfc + f) * N, g1);
            load_n<T, N>(die2 + ((ti * p.nant + a2) * p.nfc + f) * N, g2);
            if constexpr (N == 4) {
                C tmp[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) tmp[k] = acc[k];
                jones3_2x2<T>((const C(&)[4])g1, (const C(&)[4])tmp, (const C(&)[4])g2,
                              (C(&)[4])acc, false);
            } else {
                acc[0] = mulc(mul(g1[0], acc[0]), g2[0]);
            }
        }
        store_n<T, N>(out + i * N, acc);
    }
}

// ---------------------------------------------------------------------------
// fused predict with DDEs
// ---------------------------------------------------------------------------
constexpr int kK = 8;  // channels per lane (stride 32)

struct FusedParams {
    const double *lmn, *uvw, *freq;
    const void *bright;  // (nsrc,nchan,C)
    const int32_t *time_index, *ant1, *ant2;
    const void *dde1, *dde2;
    void *out;  // (nrow,nchan,C) source sum only; epilogue applied by predict_vis_kernel
    double cst;
    long long nsrc, nrow, ntime, nant;
    int nchan;
    int ncorr;  // C
};

// N = 4 with MAT: 2x2 products; otherwise N element-wise correlations
template <typename T, int N, bool MAT, bool EXACT>
__global__ void __launch_bounds__(128) fused_dde_kernel(const FusedParams p) {
    using C = Cx<T>;
    const int lane = threadIdx.x & 31;
    const int segs = (p.nchan + 32 * kK - 1) / (32 * kK);
    const long long warp_global = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const C *dde1 = (const C *)p.dde1, *dde2 = (const C *)p.dde2, *bright = (const C *)p.bright;
    C *out = (C *)p.out;
    double dnu = 0.0;
    if (!EXACT && p.nchan > 1) dnu = (p.freq[p.nchan - 1] - p.freq[0]) / (double)(p.nchan - 1);

    for (long long wi = warp_global; wi < p.nrow * segs; wi += nwarps) {
        // consecutive warps take consecutive rows of the same channel segment so that
        // concurrently running warps share DDE (time, antenna) slices in L2
        const long long seg = wi / p.nrow, r = wi - seg * p.nrow;
        const int fbase = (int)seg * 32 * kK + lane;
        const long long ti = p.time_index[r], a1 = p.ant1[r], a2 = p.ant2[r];
        const double u = p.uvw[3 * r], v = p.uvw[3 * r + 1], w = p.uvw[3 * r + 2];
        C acc[kK][N];
#pragma unroll
        for (int k = 0; k < kK; ++k)
#pragma unroll
            for (int c = 0; c < N; ++c) acc[k][c] = {T(0), T(0)};

        for (long long s = 0; s < p.nsrc; ++s) {
            const double phi = __dmul_rn(
                p.cst, phase_dot(p.lmn[3 * s], p.lmn[3 * s + 1], p.lmn[3 * s + 2], u, v, w, false));
            C2<double> z = {1.0, 0.0}, d = {1.0, 0.0};
            if (!EXACT) {
                z = cis_fast(__dmul_rn(phi, p.freq[min(fbase, p.nchan - 1)]));
                d = cis_fast(__dmul_rn(phi, 32.0 * dnu));
            }
            const C *e1p = dde1 + (((s * p.ntime + ti) * p.nant + a1) * p.nchan) * N;
            const C *e2p = dde2 + (((s * p.ntime + ti) * p.nant + a2) * p.nchan) * N;
            const C *bp = bright + (s * p.nchan) * N;
#pragma unroll
            for (int k = 0; k < kK; ++k) {
                const int f = fbase + 32 * k;
                if (f < p.nchan) {
                    if (EXACT) z = cis_fast(__dmul_rn(phi, p.freq[f]));
                    const C zz = {(T)z.re, (T)z.im};
                    C b[N], e1[N], e2[N], x[N];
                    load_n<T, N>(bp + (long long)f * N, b);
                    load_n<T, N>(e1p + (long long)f * N, e1);
                    load_n<T, N>(e2p + (long long)f * N, e2);
#pragma unroll
                    for (int c = 0; c < N; ++c) x[c] = mul(zz, b[c]);  // K * brightness
                    if constexpr (MAT) {
                        jones3_2x2<T>((const C(&)[4])e1, (const C(&)[4])x, (const C(&)[4])e2,
                                      (C(&)[4])acc[k], true);
                    } else {
#pragma unroll
                        for (int c = 0; c < N; ++c)
                            acc[k][c] = add(acc[k][c], mulc(mul(e1[c], x[c]), e2[c]));
                    }
                }
                if (!EXACT) z = cmul(z, d);
            }
        }
#pragma unroll
        for (int k = 0; k < kK; ++k) {
            const int f = fbase + 32 * k;
            if (f < p.nchan) store_n<T, N>(out + (r * p.nchan + f) * N, acc[k]);
        }
    }
}

// ---------------------------------------------------------------------------
// fused predict with DDEs, antenna-tiled (2x2 Jones, complex128, rows ordered by time)
// ---------------------------------------------------------------------------
// The gather kernel above re-reads two 64-byte Jones matrices per term (128 B/term of L2
// traffic).  Here a CTA owns (one timestep, up to 256 rows of it, FT channels); per source it
// stages E[s,t,:,f-tile] for ALL antennas once in shared memory (cp.async, one source ahead),
// precombines A_p = E_p * B_s (africanus/rime/predict.py:103-117 re-associated:
// E1 (K B) E2^H = K (E1 B) E2^H since K is a scalar), and every row then reads A_a1 and E_a2
// from shared memory: each DDE element is fetched from L2/HBM once per ~256 rows instead of
// once per row.  lane <-> row (so the phasor advances by a rotation along the thread's own
// channel run), warp <-> (row group, run of 4 channels); shared-memory matrices are stored
// [chan][antenna][64 B] with the 16-byte chunks XOR-swizzled by antenna so that lanes reading
// consecutive antennas hit distinct banks.
constexpr int kTileRowsMax = 512;  // rows per CTA = (16 / nck) * 32
constexpr int kTileCHD = 4;      // channels per thread
constexpr int kTileThreads = 512;

struct TiledParams {
    const double *lmn, *uvw, *freq;
    const double *bright;  // (nsrc,nchan,4) complex
    const int32_t *ant1, *ant2;
    const int32_t *row_start;  // (ntime+1) first row of each timestep
    const double *dde1, *dde2;
    double *out;
    double cst;
    long long nsrc, nrow, ntime, nant;
    int nchan;
    int nck;      // channel runs per CTA (1 or 2): FT = nck * 4
    int same_dde; // dde1 == dde2: stage one copy
};

__device__ __forceinline__ unsigned sw_off(int fl, int a, int na, int c) {
    // byte offset of 16-byte chunk c of the matrix of (channel fl, antenna a)
    return (unsigned)(((fl * na + a) << 6) + ((c ^ ((a >> 1) & 3)) << 4));
}

__device__ __forceinline__ void lds_mat(const unsigned char *base, int fl, int a, int na,
                                        Cx<double> (&m)[4]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const double2 v = *reinterpret_cast<const double2 *>(base + sw_off(fl, a, na, c));
        m[c].re = v.x;
        m[c].im = v.y;
    }
}

template <bool EXACT>
__global__ void __launch_bounds__(kTileThreads, 1) fused_dde_tiled_kernel(const TiledParams p) {
    using C = Cx<double>;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nck = p.nck, ft = nck * kTileCHD;
    const int na = (int)p.nant;
    const int t = blockIdx.y;
    const int f0 = blockIdx.z * ft;
    const int tile_rows = (16 / p.nck) * 32;
    const long long rbeg = p.row_start[t] + (long long)blockIdx.x * tile_rows;
    const long long rend = min((long long)p.row_start[t + 1], rbeg + tile_rows);
    if (rbeg >= rend) return;

    // shared memory: E1 slots [2], (E2 slots [2] unless same), A, B slots [2], anchors, steps
    const size_t mat_bytes = (size_t)ft * na * 64;
    unsigned char *e1s = smem;                                   // [2][mat_bytes]
    unsigned char *e2s = p.same_dde ? e1s : e1s + 2 * mat_bytes;  // [2][mat_bytes]
    unsigned char *as = e2s + 2 * mat_bytes;                      // [mat_bytes]
    unsigned char *bs = as + mat_bytes;                           // [2][ft*64]
    C2<double> *anch = reinterpret_cast<C2<double> *>(bs + 2 * (size_t)ft * 64);  // [nck][tile_rows]
    C2<double> *dstp = anch + kTileRowsMax;                                       // [tile_rows]
    double *fq = reinterpret_cast<double *>(dstp + kTileRowsMax);                 // [ft]

    // consumer role
    const int ck = warp % nck;
    const int row_local = (warp / nck) * 32 + lane;  // 16 warps / nck groups of 32 rows
    const long long r = rbeg + row_local;
    const bool row_ok = r < rend;
    int a1 = 0, a2 = 0;
    if (row_ok) {
        a1 = p.ant1[r];
        a2 = p.ant2[r];
    }
    // anchor role: thread tid < tile_rows owns row rbeg + tid
    double pu = 0, pv = 0, pw = 0;
    {
        const int arow = tid < tile_rows ? tid : tid - tile_rows;
        if (arow < tile_rows && rbeg + arow < rend) {
            pu = p.uvw[3 * (rbeg + arow)];
            pv = p.uvw[3 * (rbeg + arow) + 1];
            pw = p.uvw[3 * (rbeg + arow) + 2];
        }
    }
    double sl = p.lmn[0], sm = p.lmn[1], sn = p.lmn[2];  // coordinates of the current source
    double dnu = 0.0;
    if (p.nchan > 1) dnu = (p.freq[p.nchan - 1] - p.freq[0]) / (double)(p.nchan - 1);
    const double nu0 = p.freq[min(f0, p.nchan - 1)];
    if (EXACT)
        for (int i = tid; i < ft; i += kTileThreads) fq[i] = p.freq[min(f0 + i, p.nchan - 1)];

    C acc[kTileCHD][4];
#pragma unroll
    for (int j = 0; j < kTileCHD; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[j][c] = {0.0, 0.0};

    const int items = na * ft;  // (antenna, channel) matrices per source
    auto issue = [&](long long s) {
        if (s < p.nsrc) {
            const int slot = (int)(s & 1);
            for (int idx = tid; idx < items; idx += kTileThreads) {
                const int fl = idx / na, a = idx - fl * na;  // antenna fastest: conflict-free smem
                const int f = f0 + fl;
                const bool ok = f < p.nchan;
                const long long g = (((s * p.ntime + t) * p.nant + a) * p.nchan + (ok ? f : 0)) * 8;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(
                                     (unsigned)__cvta_generic_to_shared(e1s + slot * mat_bytes) +
                                     sw_off(fl, a, na, c)),
                                 "l"(p.dde1 + g + 2 * c), "r"(ok ? 16 : 0));
                    if (!p.same_dde)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(
                                         (unsigned)__cvta_generic_to_shared(e2s + slot * mat_bytes) +
                                         sw_off(fl, a, na, c)),
                                     "l"(p.dde2 + g + 2 * c), "r"(ok ? 16 : 0));
                }
            }
            if (tid < ft * 4) {  // brightness of this source: ft matrices of 4 chunks
                const int fl = tid >> 2, c = tid & 3;
                const int f = f0 + fl;
                const bool ok = f < p.nchan;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(
                                 (unsigned)__cvta_generic_to_shared(bs + slot * (size_t)ft * 64 + tid * 16)),
                             "l"(p.bright + ((s * p.nchan + (ok ? f : 0)) * 4 + c) * 2), "r"(ok ? 16 : 0));
            }
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };

    issue(0);
    for (long long s = 0; s < p.nsrc; ++s) {
        const int slot = (int)(s & 1);
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();  // consume(s-1) finished everywhere; E(s), B(s) landed
        issue(s + 1);
        // ---- A_p = E1_p * B_s for every (antenna, channel) of the tile
        const unsigned char *e1 = e1s + slot * mat_bytes;
        const unsigned char *e2 = e2s + slot * mat_bytes;
        for (int idx = tid; idx < items; idx += kTileThreads) {
            const int fl = idx / na, a = idx - fl * na;
            C b[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const double2 v = *reinterpret_cast<const double2 *>(bs + slot * (size_t)ft * 64 + (fl * 4 + c) * 16);
                b[c] = {v.x, v.y};
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {  // one output row of E*B at a time (register pressure)
                const double2 v0 = *reinterpret_cast<const double2 *>(e1 + sw_off(fl, a, na, 2 * h));
                const double2 v1 = *reinterpret_cast<const double2 *>(e1 + sw_off(fl, a, na, 2 * h + 1));
                const C e0 = {v0.x, v0.y}, e1m = {v1.x, v1.y};
                const C m0 = add(mul(e0, b[0]), mul(e1m, b[2]));
                const C m1 = add(mul(e0, b[1]), mul(e1m, b[3]));
                *reinterpret_cast<double2 *>(as + sw_off(fl, a, na, 2 * h)) = make_double2(m0.re, m0.im);
                *reinterpret_cast<double2 *>(as + sw_off(fl, a, na, 2 * h + 1)) = make_double2(m1.re, m1.im);
            }
        }
        // ---- phasor anchors for the rows of this CTA: thread tid < tile_rows computes the
        // first-run anchor of row tid, thread tile_rows <= tid < 2*tile_rows its channel step
        // (when the CTA has fewer than 2*tile_rows threads the first half does both)
        {
            const bool two_halves = 2 * tile_rows <= kTileThreads;
            const int arow = tid < tile_rows ? tid : tid - tile_rows;
            const bool do_z = tid < tile_rows;
            const bool do_d = two_halves ? (tid >= tile_rows && tid < 2 * tile_rows) : do_z;
            if (do_z || do_d) {
                double phi = 0.0;
                const bool live = rbeg + arow < rend;
                if (live)
                    phi = __dmul_rn(p.cst, phase_dot(sl, sm, sn, pu, pv, pw, false));
                if (EXACT) {
                    if (do_z) dstp[arow].re = phi;  // exact mode: one sincos per channel later
                } else {
                    if (do_z) anch[arow] = live ? cis_fast(__dmul_rn(phi, nu0)) : C2<double>{0.0, 0.0};
                    if (do_d) dstp[arow] = live ? cis_fast(__dmul_rn(phi, dnu)) : C2<double>{0.0, 0.0};
                }
            }
        }
        // source coordinates for the next iteration (latency hidden behind consume)
        if (s + 1 < p.nsrc) {
            sl = p.lmn[3 * (s + 1)];
            sm = p.lmn[3 * (s + 1) + 1];
            sn = p.lmn[3 * (s + 1) + 2];
        }
        __syncthreads();  // A(s), anchors(s) visible
        // ---- consume: V[r,f] += K * (A_a1 * E_a2^H)
        if (row_ok) {
            C2<double> z = anch[row_local];
            const C2<double> d = dstp[row_local];
            if (!EXACT && ck > 0) {  // run ck starts at z0 * d^(4*ck)
                C2<double> D = d;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const double re = D.re * D.re - D.im * D.im;
                    D.im = 2.0 * D.re * D.im;
                    D.re = re;
                }
                for (int q = 0; q < ck; ++q) z = cmul(z, D);
            }
#pragma unroll
            for (int j = 0; j < kTileCHD; ++j) {
                const int fl = ck * kTileCHD + j;
                if (EXACT) z = cis_fast(__dmul_rn(d.re, fq[fl]));
                C em[4];
                lds_mat(e2, fl, a2, na, em);
                const C zz = {z.re, z.im};
#pragma unroll
                for (int h = 0; h < 2; ++h) {  // output rows of the 2x2 product, one at a time
                    C a0, a1m;
                    {
                        const double2 v0 = *reinterpret_cast<const double2 *>(as + sw_off(fl, a1, na, 2 * h));
                        const double2 v1 = *reinterpret_cast<const double2 *>(as + sw_off(fl, a1, na, 2 * h + 1));
                        a0 = {v0.x, v0.y};
                        a1m = {v1.x, v1.y};
                    }
                    const C m0 = add(mulc(a0, em[0]), mulc(a1m, em[1]));
                    const C m1 = add(mulc(a0, em[2]), mulc(a1m, em[3]));
                    acc[j][2 * h] = add(acc[j][2 * h], mul(zz, m0));
                    acc[j][2 * h + 1] = add(acc[j][2 * h + 1], mul(zz, m1));
                }
                if (!EXACT) z = cmul(z, d);
            }
        }
    }
    if (row_ok) {
#pragma unroll
        for (int j = 0; j < kTileCHD; ++j) {
            const int f = f0 + ck * kTileCHD + j;
            if (f < p.nchan) store_n<double, 4>(reinterpret_cast<C *>(p.out) + (r * p.nchan + f) * 4, acc[j]);
        }
    }
}

// row_start[t] = first row with time_index >= t (rows sorted by time); flags[0] |= 1 when
// the rows are not sorted; flags[1] = max rows in one timestep
__global__ void time_ranges_kernel(const int32_t *time_index, long long nrow, long long ntime,
                                   int32_t *row_start, int *flags) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i + 1 < nrow && time_index[i + 1] < time_index[i]) atomicOr(&flags[0], 1);
    if (i < nrow && (time_index[i] < 0 || time_index[i] >= ntime)) atomicOr(&flags[0], 2);
    if (i <= ntime) {
        long long lo = 0, hi = nrow;  // lower_bound(time_index, i)
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (time_index[mid] < i) lo = mid + 1; else hi = mid;
        }
        row_start[i] = (int32_t)lo;
    }
}
__global__ void max_rows_kernel(const int32_t *row_start, long long ntime, int *flags) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t < ntime) atomicMax(&flags[1], row_start[t + 1] - row_start[t]);
}

int grid_for(long long total, int threads) {
    long long blocks = (total + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <typename T>
int launch_predict(const PredictParams &p, int jones_mode, cudaStream_t stream) {
    const long long total = p.nrow * p.nfc;
    if (total <= 0) return 0;
    if (jones_mode == AFR_JONES_2X2)
        predict_vis_kernel<T, 4><<<grid_for(total, 256), 256, 0, stream>>>(p);
    else
        predict_vis_kernel<T, 1><<<grid_for(total, 256), 256, 0, stream>>>(p);
    AFR_LAUNCH_OK();
    return 0;
}

template <typename T>
int launch_fused_dde(const FusedParams &p, int jones_mode, bool exact, cudaStream_t stream) {
    const int segs = (p.nchan + 32 * kK - 1) / (32 * kK);
    const long long warps = p.nrow * segs;
    if (warps <= 0) return 0;
    const int grid = grid_for(warps * 32, 128);
#define AFR_LAUNCH_FUSED(N, MAT)                                                        \
    do {                                                                                \
        if (exact)                                                                      \
            fused_dde_kernel<T, N, MAT, true><<<grid, 128, 0, stream>>>(p);             \
        else                                                                            \
            fused_dde_kernel<T, N, MAT, false><<<grid, 128, 0, stream>>>(p);            \
    } while (0)
    if (jones_mode == AFR_JONES_2X2) {
        AFR_LAUNCH_FUSED(4, true);
    } else if (p.ncorr == 1) {
        AFR_LAUNCH_FUSED(1, false);
    } else if (p.ncorr == 2) {
        AFR_LAUNCH_FUSED(2, false);
    } else if (p.ncorr == 4) {
        AFR_LAUNCH_FUSED(4, false);
    } else {
        return fail("afr_predict_fused: diagonal Jones with DDEs supports ncorr in (1, 2, 4)");
    }
#undef AFR_LAUNCH_FUSED
    AFR_LAUNCH_OK();
    return 0;
}

}  // namespace
}  // namespace afr

using namespace afr;

extern "C" int afr_predict_vis(const int32_t *time_index, const int32_t *antenna1,
                               const int32_t *antenna2, const void *dde1, const void *source_coh,
                               const void *dde2, const void *die1, const void *base_vis,
                               const void *die2, int64_t nsrc, int64_t nrow, int64_t ntime,
                               int64_t nant, int64_t nchan, int64_t ncorr, int jones_mode,
                               int is_c64, void *out, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    // africanus/rime/predict.py:403-407
    AFR_REQUIRE((dde1 == nullptr) == (dde2 == nullptr),
                "Both dde1_jones and dde2_jones must be present or absent");
    AFR_REQUIRE((die1 == nullptr) == (die2 == nullptr),
                "Both die1_jones and die2_jones must be present or absent");
    AFR_REQUIRE(dde1 || source_coh || die1 || base_vis, "No Jones Matrices were supplied");
    AFR_REQUIRE(jones_mode == AFR_JONES_DIAG || (jones_mode == AFR_JONES_2X2 && ncorr == 4),
                "Jones Matrix Correlations were mismatched");
    AFR_REQUIRE(nsrc >= 0 && nrow >= 0 && nchan >= 0 && ncorr >= 1, "bad extent");
    PredictParams p{};
    p.time_index = time_index;
    p.ant1 = antenna1;
    p.ant2 = antenna2;
    p.dde1 = dde1;
    p.coh = source_coh;
    p.dde2 = dde2;
    p.die1 = die1;
    p.bvis = base_vis;
    p.die2 = die2;
    p.acc_init = nullptr;
    p.out = out;
    p.nsrc = nsrc;
    p.nrow = nrow;
    p.ntime = ntime;
    p.nant = nant;
    p.nfc = jones_mode == AFR_JONES_2X2 ? nchan : nchan * ncorr;
    return is_c64 ? launch_predict<float>(p, jones_mode, stream)
                  : launch_predict<double>(p, jones_mode, stream);
}

extern "C" int afr_predict_fused(const double *lm, const double *uvw, const double *freq,
                                 const void *brightness, const int32_t *time_index,
                                 const int32_t *antenna1, const int32_t *antenna2,
                                 const void *dde1, const void *dde2, const void *die1,
                                 const void *base_vis, const void *die2, int64_t nsrc,
                                 int64_t nrow, int64_t ntime, int64_t nant, int64_t nchan,
                                 int64_t ncorr, int jones_mode, int convention, int chan_mode,
                                 int out_c64, void *out, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AFR_REQUIRE(convention == AFR_FOURIER || convention == AFR_CASA,
                "convention not in ('fourier', 'casa')");
    AFR_REQUIRE((dde1 == nullptr) == (dde2 == nullptr),
                "Both dde1_jones and dde2_jones must be present or absent");
    AFR_REQUIRE((die1 == nullptr) == (die2 == nullptr),
                "Both die1_jones and die2_jones must be present or absent");
    AFR_REQUIRE(jones_mode == AFR_JONES_DIAG || (jones_mode == AFR_JONES_2X2 && ncorr == 4),
                "Jones Matrix Correlations were mismatched");
    AFR_REQUIRE(nsrc >= 0 && nrow >= 0 && nchan >= 0 && ncorr >= 1 && nchan < (1LL << 30),
                "bad extent");
    if (nrow == 0 || nchan == 0) return 0;
    const bool exact = chan_mode == AFR_CHAN_EXACT;
    // rime/phase.py:29-34
    const double cst = convention == AFR_FOURIER ? -kTwoPiOverC : kTwoPiOverC;

    Scratch lmn;
    AFR_CUDA_OK(lmn.alloc(sizeof(double) * 3 * (size_t)nsrc, stream));
    int rc = launch_lm_to_lmn(lm, nsrc, kLmnPhaseClamp, false, (double *)lmn.ptr, stream);
    if (rc) return rc;

    const size_t elem = out_c64 ? 8 : 16;
    if (nsrc == 0) {
        AFR_CUDA_OK(cudaMemsetAsync(out, 0, (size_t)nrow * nchan * ncorr * elem, stream));
    } else if (dde1 == nullptr) {
        // point-source sum: phasor-stream kernel with the brightness as complex "image"
        // (complex64 brightness feeds the FP32 accumulator variant directly)
        rc = run_phasor_stream(uvw, nrow, (const double *)lmn.ptr, nsrc, brightness, true, nullptr,
                               freq, nchan, ncorr, cst, false, /*adjoint=*/false, exact,
                               out_c64 != 0, out, stream);
        if (rc) return rc;
        note_fused_path(AFR_PATH_POINT);
    } else {
        FusedParams f{};
        f.lmn = (const double *)lmn.ptr;
        f.uvw = uvw;
        f.freq = freq;
        f.bright = brightness;
        f.time_index = time_index;
        f.ant1 = antenna1;
        f.ant2 = antenna2;
        f.dde1 = dde1;
        f.dde2 = dde2;
        f.out = out;
        f.cst = cst;
        f.nsrc = nsrc;
        f.nrow = nrow;
        f.ntime = ntime;
        f.nant = nant;
        f.nchan = (int)nchan;
        f.ncorr = (int)ncorr;
        bool done = false;
        if (!out_c64 && jones_mode == AFR_JONES_2X2 && nsrc > 0) {
            // antenna-tiled kernel: needs rows ordered by time and the tile in shared memory
            Scratch rs, fl;
            AFR_CUDA_OK(rs.alloc(sizeof(int32_t) * (size_t)(ntime + 1), stream));
            AFR_CUDA_OK(fl.alloc(sizeof(int) * 2, stream));
            AFR_CUDA_OK(cudaMemsetAsync(fl.ptr, 0, sizeof(int) * 2, stream));
            const long long n = std::max<long long>(nrow, ntime + 1);
            time_ranges_kernel<<<(int)((n + 255) / 256), 256, 0, stream>>>(
                time_index, nrow, ntime, (int32_t *)rs.ptr, (int *)fl.ptr);
            AFR_LAUNCH_OK();
            max_rows_kernel<<<(int)((ntime + 255) / 256), 256, 0, stream>>>((const int32_t *)rs.ptr,
                                                                            ntime, (int *)fl.ptr);
            AFR_LAUNCH_OK();
            // warp-specialised kernel (afr_rime_ws.cu): TMA needs 16-byte aligned sources and
            // the three-stage antenna tile must fit in shared memory
            // (the 2048-row x 1-channel antenna-mode tile is 3.4x smaller per antenna than the
            // 512-row x 4-channel one, so SKA-size arrays still fit in antenna mode)
            const bool fit4 = dde_ws_row_tile_channels(nant) > 0;  // 4, 2 or 1 channels per 512-row tile
            const bool fit1 = dde_ws_smem_bytes(nant, 1, true) <= 220 * 1024;
            const bool ws_ok = (fit4 || fit1) &&
                               reinterpret_cast<uintptr_t>(dde1) % 16 == 0 &&
                               reinterpret_cast<uintptr_t>(dde2) % 16 == 0 &&
                               reinterpret_cast<uintptr_t>(brightness) % 16 == 0 &&
                               !(getenv("AFR_DDE_WS") && atoi(getenv("AFR_DDE_WS")) == 0);
            Scratch antuvw, antok, rowmap, used4, dup;
            // (time, antenna1, antenna2) -> row map of the GEMM path (afr_rime_mma.cu); AFR_DDE_MMA=0
            // keeps the scalar warp-specialised kernel
            const int n4 = (int)((nant + 3) / 4);
            const bool mma_ok = ws_ok && nant <= 1024 && nrow < (1LL << 31) &&
                                (double)ntime * nant * nant * 4.0 <= 1.0e9 &&
                                !(getenv("AFR_DDE_MMA") && atoi(getenv("AFR_DDE_MMA")) == 0);
            if (mma_ok) {
                AFR_CUDA_OK(rowmap.alloc(sizeof(int32_t) * (size_t)(ntime * nant * nant), stream));
                AFR_CUDA_OK(used4.alloc((size_t)n4 * n4, stream));
                AFR_CUDA_OK(dup.alloc(sizeof(int), stream));
                rc = launch_baseline_map(time_index, antenna1, antenna2, nrow, ntime, nant,
                                         (int32_t *)rowmap.ptr, (uint8_t *)used4.ptr, (int *)dup.ptr, stream);
                if (rc) return rc;
            }
            if (ws_ok) {
                AFR_CUDA_OK(antuvw.alloc(sizeof(double) * 3 * (size_t)(ntime * nant), stream));
                AFR_CUDA_OK(antok.alloc(sizeof(int), stream));
                const int one = 1;
                AFR_CUDA_OK(cudaMemcpyAsync(antok.ptr, &one, sizeof(int), cudaMemcpyHostToDevice, stream));
                rc = launch_antenna_uvw(uvw, antenna1, antenna2, (const int32_t *)rs.ptr, ntime, nant,
                                        (const double *)lmn.ptr, nsrc, freq, nchan, cst,
                                        (double *)antuvw.ptr, (int *)antok.ptr, stream);
                if (rc) return rc;
            }
            int hflags[2] = {0, 0};
            int hant = 0;
            AFR_CUDA_OK(cudaMemcpyAsync(hflags, fl.ptr, sizeof(hflags), cudaMemcpyDeviceToHost, stream));
            if (ws_ok)
                AFR_CUDA_OK(cudaMemcpyAsync(&hant, antok.ptr, sizeof(int), cudaMemcpyDeviceToHost, stream));
            int hdup = 1;
            std::vector<uint8_t> hused;
            if (mma_ok) {
                hused.resize((size_t)n4 * n4);
                AFR_CUDA_OK(cudaMemcpyAsync(&hdup, dup.ptr, sizeof(int), cudaMemcpyDeviceToHost, stream));
                AFR_CUDA_OK(cudaMemcpyAsync(hused.data(), used4.ptr, hused.size(), cudaMemcpyDeviceToHost, stream));
            }
            AFR_CUDA_OK(cudaStreamSynchronize(stream));
            const bool same = dde1 == dde2;
            const char *am_env = getenv("AFR_DDE_ANT");  // 0 forces the per-row phasor mode
            const bool ant_mode = hant != 0 && !(am_env && atoi(am_env) == 0);
            if (mma_ok && ant_mode && hdup == 0 && hflags[0] == 0 && hflags[1] > 0) {
                // every baseline (t, a1, a2) names one row: source sum as a GEMM per (time, channel)
                DdeMmaParams mp{};
                mp.lmn = (const double *)lmn.ptr;
                mp.freq = freq;
                mp.bright = (const double *)brightness;
                mp.dde1 = (const double *)dde1;
                mp.dde2 = (const double *)dde2;
                mp.ant_uvw = (const double *)antuvw.ptr;
                mp.rowmap = (const int32_t *)rowmap.ptr;
                mp.out = (double *)out;
                mp.cst = cst;
                mp.nsrc = nsrc;
                mp.ntime = ntime;
                mp.nant = nant;
                mp.nchan = (int)nchan;
                mp.same_dde = same ? 1 : 0;
                rc = launch_fused_dde_mma(mp, dde_mma_passes(hused, nant), stream);
                if (rc) return rc;
                note_fused_path(AFR_PATH_DDE_MMA_ANT);
                done = true;
            }
            const bool ws_fits = !done && (fit4 || (fit1 && ant_mode && hflags[1] > 512));
            if (ws_ok && ws_fits && hflags[0] == 0 && hflags[1] > 0 && nrow < (1LL << 31) && nant <= 1024) {
                Scratch perm;
                AFR_CUDA_OK(perm.alloc(sizeof(int32_t) * (size_t)nrow, stream));
                // antenna mode with 2048-row tiles pairs baselines that share antenna 1
                rc = launch_row_tile_order(time_index, antenna1, antenna2, nrow, ntime,
                                           ant_mode && hflags[1] > 512, (int32_t *)perm.ptr, stream);
                if (rc) return rc;
                DdeWsParams wp{};
                wp.perm = (const int32_t *)perm.ptr;
                wp.lmn = (const double *)lmn.ptr;
                wp.uvw = uvw;
                wp.freq = freq;
                wp.bright = (const double *)brightness;
                wp.ant1 = antenna1;
                wp.ant2 = antenna2;
                wp.row_start = (const int32_t *)rs.ptr;
                wp.dde1 = (const double *)dde1;
                wp.dde2 = (const double *)dde2;
                wp.ant_uvw = (const double *)antuvw.ptr;
                wp.out = (double *)out;
                wp.cst = cst;
                wp.nsrc = nsrc;
                wp.nrow = nrow;
                wp.ntime = ntime;
                wp.nant = nant;
                wp.nchan = (int)nchan;
                wp.same_dde = same ? 1 : 0;
                rc = launch_fused_dde_ws(wp, hflags[1], exact, ant_mode, stream);
                if (rc) return rc;
                note_fused_path(ant_mode ? AFR_PATH_DDE_WS_ANT : AFR_PATH_DDE_WS_ROW);
                done = true;
            }
            int nck = 2;
            auto smem_for = [&](int k) {
                const size_t mat = (size_t)k * kTileCHD * nant * 64;
                return (same ? 3 : 5) * mat + 2 * (size_t)k * kTileCHD * 64 +
                       2 * (size_t)kTileRowsMax * 16 + (size_t)k * kTileCHD * 8;
            };
            if (smem_for(nck) > 200 * 1024) nck = 1;
            if (!done && hflags[0] == 0 && hflags[1] > 0 && smem_for(nck) <= 200 * 1024) {
                TiledParams tp{};
                tp.lmn = (const double *)lmn.ptr;
                tp.uvw = uvw;
                tp.freq = freq;
                tp.bright = (const double *)brightness;
                tp.ant1 = antenna1;
                tp.ant2 = antenna2;
                tp.row_start = (const int32_t *)rs.ptr;
                tp.dde1 = (const double *)dde1;
                tp.dde2 = (const double *)dde2;
                tp.out = (double *)out;
                tp.cst = cst;
                tp.nsrc = nsrc;
                tp.nrow = nrow;
                tp.ntime = ntime;
                tp.nant = nant;
                tp.nchan = (int)nchan;
                tp.nck = nck;
                tp.same_dde = same ? 1 : 0;
                const int tile_rows = (16 / nck) * 32;
                const int ft = nck * kTileCHD;
                dim3 grid((unsigned)((hflags[1] + tile_rows - 1) / tile_rows), (unsigned)ntime,
                          (unsigned)((nchan + ft - 1) / ft));
                AFR_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "afr_predict_fused: grid too large");
                const size_t smem = smem_for(nck);
                // rows not covered by any timestep range cannot exist once the order check passed
                if (exact) {
                    AFR_CUDA_OK(cudaFuncSetAttribute(fused_dde_tiled_kernel<true>,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    fused_dde_tiled_kernel<true><<<grid, kTileThreads, smem, stream>>>(tp);
                } else {
                    AFR_CUDA_OK(cudaFuncSetAttribute(fused_dde_tiled_kernel<false>,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    fused_dde_tiled_kernel<false><<<grid, kTileThreads, smem, stream>>>(tp);
                }
                AFR_LAUNCH_OK();
                note_fused_path(AFR_PATH_DDE_TILED);
                done = true;
            }
        }
        if (!done) note_fused_path(AFR_PATH_DDE_GATHER);
        if (!done)
            rc = out_c64 ? launch_fused_dde<float>(f, jones_mode, exact, stream)
                         : launch_fused_dde<double>(f, jones_mode, exact, stream);
        if (rc) return rc;
    }

    if (base_vis != nullptr || die1 != nullptr) {
        // epilogue: out = G1 (out + B) G2^H  (predict.py:329-373), in place
        PredictParams p{};
        p.time_index = time_index;
        p.ant1 = antenna1;
        p.ant2 = antenna2;
        p.die1 = die1;
        p.bvis = base_vis;
        p.die2 = die2;
        p.acc_init = out;
        p.out = out;
        p.nsrc = 0;
        p.nrow = nrow;
        p.ntime = ntime;
        p.nant = nant;
        p.nfc = jones_mode == AFR_JONES_2X2 ? nchan : nchan * ncorr;
        rc = out_c64 ? launch_predict<float>(p, jones_mode, stream)
                     : launch_predict<double>(p, jones_mode, stream);
        if (rc) return rc;
    }
    return 0;
}

// Full-RIME predict with the DDE Jones sampled from the plane-reduced beam INSIDE the predict kernel
// (SURVEY 8f-1 proper; reference: experimental/rime/fused/terms/cube_dde.py:96-313 samples the cube in
// its fused loop): antenna-phasor mode of the warp-specialised kernel, whose producers form each
// antenna's Jones from the two frequency planes of the channel (afr_beam_plane_reduce) before they
// precombine it with the brightness -- the (source,time,ant,chan,2,2) array never exists.
// used[0] (host) = 1 when the predict ran; 0 when this path does not apply (uvw that are not
// differences of antenna coordinates within the admission bound, rows not ordered by time, antenna
// tile too large for shared memory): the caller then takes the chunked route, nothing was written.
extern "C" int afr_predict_fused_planes(const double *lm, const double *uvw, const double *freq,
                                        const void *brightness, const int32_t *time_index,
                                        const int32_t *antenna1, const int32_t *antenna2, const double *planes,
                                        const double *fd, int64_t nud, const void *feed_rotation,
                                        const void *die1, const void *base_vis,
                                        const void *die2, int64_t nsrc, int64_t nrow, int64_t ntime,
                                        int64_t nant, int64_t nchan, int convention, int *used, void *out,
                                        void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AFR_REQUIRE(used != nullptr, "afr_predict_fused_planes: used must not be NULL");
    *used = 0;
    AFR_REQUIRE(convention == AFR_FOURIER || convention == AFR_CASA, "convention not in ('fourier', 'casa')");
    AFR_REQUIRE((die1 == nullptr) == (die2 == nullptr),
                "Both die1_jones and die2_jones must be present or absent");
    AFR_REQUIRE(nsrc >= 1 && nrow >= 1 && nchan >= 1 && ntime >= 1 && nant >= 1 && nud >= 2 &&
                    nchan < (1LL << 30) && nrow < (1LL << 31) && nant <= 1024,
                "afr_predict_fused_planes: bad extent");
    if (reinterpret_cast<uintptr_t>(planes) % 16 != 0 || reinterpret_cast<uintptr_t>(brightness) % 16 != 0 ||
        reinterpret_cast<uintptr_t>(feed_rotation) % 16 != 0 ||
        dde_ws_smem_bytes(nant, 1, true, true) > 220 * 1024)
        return 0;
    const double cst = convention == AFR_FOURIER ? -kTwoPiOverC : kTwoPiOverC;
    Scratch lmn, rs, fl, antuvw, antok, perm;
    AFR_CUDA_OK(lmn.alloc(sizeof(double) * 3 * (size_t)nsrc, stream));
    int rc = launch_lm_to_lmn(lm, nsrc, kLmnPhaseClamp, false, (double *)lmn.ptr, stream);
    if (rc) return rc;
    AFR_CUDA_OK(rs.alloc(sizeof(int32_t) * (size_t)(ntime + 1), stream));
    AFR_CUDA_OK(fl.alloc(sizeof(int) * 2, stream));
    AFR_CUDA_OK(cudaMemsetAsync(fl.ptr, 0, sizeof(int) * 2, stream));
    const long long n = std::max<long long>(nrow, ntime + 1);
    time_ranges_kernel<<<(int)((n + 255) / 256), 256, 0, stream>>>(time_index, nrow, ntime, (int32_t *)rs.ptr,
                                                                   (int *)fl.ptr);
    AFR_LAUNCH_OK();
    max_rows_kernel<<<(int)((ntime + 255) / 256), 256, 0, stream>>>((const int32_t *)rs.ptr, ntime, (int *)fl.ptr);
    AFR_LAUNCH_OK();
    AFR_CUDA_OK(antuvw.alloc(sizeof(double) * 3 * (size_t)(ntime * nant), stream));
    AFR_CUDA_OK(antok.alloc(sizeof(int), stream));
    const int one = 1;
    AFR_CUDA_OK(cudaMemcpyAsync(antok.ptr, &one, sizeof(int), cudaMemcpyHostToDevice, stream));
    rc = launch_antenna_uvw(uvw, antenna1, antenna2, (const int32_t *)rs.ptr, ntime, nant, (const double *)lmn.ptr,
                            nsrc, freq, nchan, cst, (double *)antuvw.ptr, (int *)antok.ptr, stream);
    if (rc) return rc;
    int hflags[2] = {0, 0}, hant = 0;
    AFR_CUDA_OK(cudaMemcpyAsync(hflags, fl.ptr, sizeof(hflags), cudaMemcpyDeviceToHost, stream));
    AFR_CUDA_OK(cudaMemcpyAsync(&hant, antok.ptr, sizeof(int), cudaMemcpyDeviceToHost, stream));
    AFR_CUDA_OK(cudaStreamSynchronize(stream));
    const char *am_env = getenv("AFR_DDE_ANT");
    if (hant == 0 || hflags[0] != 0 || hflags[1] <= 0 || (am_env && atoi(am_env) == 0)) return 0;
    AFR_CUDA_OK(perm.alloc(sizeof(int32_t) * (size_t)nrow, stream));
    rc = launch_row_tile_order(time_index, antenna1, antenna2, nrow, ntime, hflags[1] > 512, (int32_t *)perm.ptr,
                               stream);
    if (rc) return rc;
    DdeWsParams wp{};
    wp.perm = (const int32_t *)perm.ptr;
    wp.lmn = (const double *)lmn.ptr;
    wp.uvw = uvw;
    wp.freq = freq;
    wp.bright = (const double *)brightness;
    wp.ant1 = antenna1;
    wp.ant2 = antenna2;
    wp.row_start = (const int32_t *)rs.ptr;
    wp.ant_uvw = (const double *)antuvw.ptr;
    wp.out = (double *)out;
    wp.cst = cst;
    wp.nsrc = nsrc;
    wp.nrow = nrow;
    wp.ntime = ntime;
    wp.nant = nant;
    wp.nchan = (int)nchan;
    wp.same_dde = 1;
    wp.planes = planes;
    wp.fd = fd;
    wp.feed = (const double *)feed_rotation;
    wp.nud = (int)nud;
    rc = launch_fused_dde_ws(wp, hflags[1], false, true, stream);
    if (rc) return rc;
    note_fused_path(AFR_PATH_DDE_WS_ANT_SAMPLED);
    if (base_vis != nullptr || die1 != nullptr) {
        PredictParams p{};
        p.time_index = time_index;
        p.ant1 = antenna1;
        p.ant2 = antenna2;
        p.die1 = die1;
        p.bvis = base_vis;
        p.die2 = die2;
        p.acc_init = out;
        p.out = out;
        p.nsrc = 0;
        p.nrow = nrow;
        p.ntime = ntime;
        p.nant = nant;
        p.nfc = nchan;
        rc = launch_predict<double>(p, AFR_JONES_2X2, stream);
        if (rc) return rc;
    }
    *used = 1;
    return 0;
}
