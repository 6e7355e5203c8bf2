// pxb_stencil3d_dense.cuh -- DENSE (full-rank) K x K x K stencil ('constant' boundaries) in ONE pass over HBM.
//
// The reference takes any dense 3-D kernel and evaluates it sample by sample over the padded array (stencil.py:356-461,
// _stencil.py:232-305: one thread per output sample, K^3 loads each).  A dense kernel has no factors to split, so the work is
// K^3 FMAs per voxel (343 for a measured 7x7x7 PSF) against 8 B of HBM traffic: the kernel is FMA-bound and the design is about
// feeding the FMA pipe, not about bytes:
//   * a CTA owns a tile of TY x TX samples and MARCHES along the slowest axis.  Per INPUT plane it stages the tile's window
//     ((TY + K - 1) x (TX + K - 1) samples, zeros outside the image = 'constant' Pad, pad.py:252-258) in shared memory, and
//     every thread SCATTERS that plane into the K output planes it contributes to: the accumulators of those K output planes
//     (K x R x VEC per thread) live in registers and shift by one slot per plane; the slot that just received its last
//     contribution is the finished output plane.  Each staged sample is therefore read from shared memory once per
//     K*K*K-tap, not once per output plane: per thread and plane (R + K - 1) rows of NV vector loads feed R*VEC*K^3 FMAs
//     (K = 7, fp32: 42 LDS.128 for 2744 FFMA);
//   * the K^3 coefficients are a by-value kernel parameter and every loop over taps is unrolled, so each FMA takes its
//     coefficient as a constant-bank operand: no coefficient loads, no index arithmetic in the tap loops;
//   * the next plane's window is fetched into registers before the accumulation and stored to the other buffer after it
//     (one barrier per plane).  Plain coalesced loads with bounds tests instead of TMA boxes: ~12 loads per thread against
//     2744 FMAs, and no 16-byte constraint on the window's first column (any centre along the rows).
// Epilogue: out = alpha * S(in) + beta * add[i % period].  z-slab cuts read the neighbour's ghost planes (as pxb_stencil3d).
// HBM traffic: 8 B/voxel (fp32) + (K - 1)/chunk for the planes two chunks share; halo re-reads are L2 hits.
// Kernels that are not cubes of 3, 5 or 7 taps are embedded in the next such cube (zero taps) by the launcher when that at most
// doubles the tap count.  Bodies are __host__ __device__ (tests/emu replays them CTA by CTA).
#pragma once
#include "pxb_tv_fast.cuh"

#if defined(__CUDACC__)
#define PXB_UNROLL _Pragma("unroll")
#define PXB_NOUNROLL _Pragma("unroll 1")
#else
#define PXB_UNROLL
#define PXB_NOUNROLL
#endif

template <class T, int K>
struct PxbD3Cfg {
    static constexpr int VEC = 16 / (int)sizeof(T);
    static constexpr int TXL = 32, TX = TXL * VEC, TY = 16, R = 2, NT = 256;
    static constexpr int BH = TY + K - 1;                     // rows of the staged window
    static constexpr int NV = (VEC + K - 1 + VEC - 1) / VEC;  // vectors covering one thread's VEC + K - 1 samples of a row
    static constexpr int PITCH = TX + (NV - 1) * VEC;         // columns of the staged window (>= TX + K - 1, multiple of VEC)
    static constexpr int NROW = (BH + 7) / 8, NCOL = (PITCH + 31) / 32;  // fetch: warp w takes rows w + 8i, lane l columns l + 32j
    static constexpr int BOX = (BH * PITCH + 31) / 32 * 32;   // elements of one staging buffer
    static constexpr int CROW = (K * K + VEC - 1) / VEC * VEC;  // coefficients of one kernel row, padded to whole 16-byte vectors
};

template <class T, int K>
struct PxbD3P {            // by-value kernel parameter
    int n0, n1, n2;        // owned planes, rows, columns
    int64_t batch, vol;    // batch items, elements between them (plane_alloc * n1 * n2)
    int c0, c1, c2;        // the kernel's entry on the output sample
    int lo_planes, hi_planes;  // readable ghost planes below plane 0 / above plane n0-1
    int chunk, nchunk;     // output planes per CTA along the marching axis
    int ntx, nty;
    T alpha, beta;
    const T* add;          // dense (batch, n0, n1, n2), nullable
    int64_t add_period;    // 0: as long as the output
    alignas(16) T coef[K * PxbD3Cfg<T, K>::CROW];  // [axis 1][axis 0][axis 2]: the K*K coefficients of one kernel row are contiguous and
                           // start on a 16-byte boundary (vector loads from the parameter bank, pxb_d3_accum)
};

// window of input plane `plane` (pointer to its sample (0, 0)) for the tile at (y0, x0) -> registers; zeros outside the image
template <class T, int K>
PXB_HD void pxb_d3_fetch(const PxbD3P<T, K>& p, const T* __restrict__ plane, int y0, int x0, int tid, T* pre) {
    using C = PxbD3Cfg<T, K>;
    const int w = tid >> 5, l = tid & 31;
    for (int i = 0; i < C::NROW; ++i) {
        const int row = w + 8 * i, y = y0 - p.c1 + row;
        const bool rok = row < C::BH && y >= 0 && y < p.n1;
        for (int j = 0; j < C::NCOL; ++j) {
            const int col = l + 32 * j, x = x0 - p.c2 + col;
            const bool ok = rok && col < C::PITCH && x >= 0 && x < p.n2;
            pre[i * C::NCOL + j] = ok ? plane[(int64_t)y * p.n2 + x] : T(0);
        }
    }
}

template <class T, int K>
PXB_HD void pxb_d3_stash(const T* pre, T* __restrict__ box, int tid) {
    using C = PxbD3Cfg<T, K>;
    const int w = tid >> 5, l = tid & 31;
    for (int i = 0; i < C::NROW; ++i) {
        const int row = w + 8 * i;
        if (row >= C::BH) break;
        for (int j = 0; j < C::NCOL; ++j) {
            const int col = l + 32 * j;
            if (col < C::PITCH) box[row * C::PITCH + col] = pre[i * C::NCOL + j];
        }
    }
}

// one kernel row (bp) of one staged input plane: window rows bp + r for the R output rows of the thread, every tap along the
// columns and every kernel plane a.  `cf`: the K*K coefficients k[a][bp][c] of this kernel row, laid out [a][c].
template <class T, int K>
PXB_HD void pxb_d3_row(const T* __restrict__ cf, const T* __restrict__ box, int yl, int xl, int bp, T (*acc)[PxbD3Cfg<T, K>::R][PxbD3Cfg<T, K>::VEC]) {
    using C = PxbD3Cfg<T, K>;
    constexpr int VEC = C::VEC;
PXB_UNROLL
    for (int r = 0; r < C::R; ++r) {
        T v[C::NV * VEC];
        const T* __restrict__ src = box + (yl + bp + r) * C::PITCH + xl;
PXB_UNROLL
        for (int n = 0; n < C::NV; ++n) {
            const PxbVec<T, VEC> t = pxb_vload<T, VEC>(src + n * VEC);
PXB_UNROLL
            for (int j = 0; j < VEC; ++j) v[n * VEC + j] = t.v[j];
        }
PXB_UNROLL
        for (int c = 0; c < K; ++c) {
            PXB_UNROLL
            for (int j = 0; j < VEC; ++j) {
                PXB_UNROLL
                for (int a = 0; a < K; ++a) acc[a][r][j] += cf[a * K + c] * v[c + j];
            }
        }
    }
}

// the same for the planes at the two ends of a chunk, where only the slots a_lo..a_hi belong to output planes of the chunk (the
// others would collect sums nobody stores: (K - 1)/chunk of all FMAs): kernel plane outermost, one uniform test per slot
template <class T, int K>
PXB_HD void pxb_d3_row_some(const T* __restrict__ cf, const T* __restrict__ box, int yl, int xl, int bp, T (*acc)[PxbD3Cfg<T, K>::R][PxbD3Cfg<T, K>::VEC],
                            int a_lo, int a_hi) {
    using C = PxbD3Cfg<T, K>;
    constexpr int VEC = C::VEC;
PXB_UNROLL
    for (int r = 0; r < C::R; ++r) {
        T v[C::NV * VEC];
        const T* __restrict__ src = box + (yl + bp + r) * C::PITCH + xl;
PXB_UNROLL
        for (int n = 0; n < C::NV; ++n) {
            const PxbVec<T, VEC> t = pxb_vload<T, VEC>(src + n * VEC);
PXB_UNROLL
            for (int j = 0; j < VEC; ++j) v[n * VEC + j] = t.v[j];
        }
PXB_UNROLL
        for (int a = 0; a < K; ++a) {
            if (a < a_lo || a > a_hi) continue;
            PXB_UNROLL
            for (int c = 0; c < K; ++c) {
                PXB_UNROLL
                for (int j = 0; j < VEC; ++j) acc[a][r][j] += cf[a * K + c] * v[c + j];
            }
        }
    }
}

// one staged input plane scattered into the K output planes it contributes to: slot a holds the output plane that takes this
// input plane with the kernel's plane a (output q = input plane - a + c0).  `coef` is laid out [kernel row][kernel plane][column].
// The loop over the kernel's rows stays ROLLED for K = 7: fully unrolled, the 7x7x7 body is 2744 FFMA = 46 KB of code, more than
// the 32 KB instruction cache an SM's warps share -- ncu on the first version: "no instruction" was the top stall (1.6 per
// issued instruction, 78 % issue-active; rolled: 0.08 and 93 %, 4.20 -> 3.60 ms); one kernel row is 392 FFMA + 6 LDS + the row's
// 49 coefficients (uniform loads from the parameter bank at a run-time row offset).  5x5x5 (1000 FFMA, 16 KB) fits and is faster
// unrolled (1.59 against 1.9 ms on 256x1024^2).
template <class T, int K>
PXB_HD void pxb_d3_accum(const T* __restrict__ coef, const T* __restrict__ box, int yl, int xl, T (*acc)[PxbD3Cfg<T, K>::R][PxbD3Cfg<T, K>::VEC]) {
    if (K <= 5) {
PXB_UNROLL
        for (int bp = 0; bp < K; ++bp) pxb_d3_row<T, K>(coef + bp * PxbD3Cfg<T, K>::CROW, box, yl, xl, bp, acc);
    } else {
PXB_NOUNROLL
        for (int bp = 0; bp < K; ++bp) pxb_d3_row<T, K>(coef + bp * PxbD3Cfg<T, K>::CROW, box, yl, xl, bp, acc);
    }
}
template <class T, int K>
PXB_HD void pxb_d3_accum_some(const T* __restrict__ coef, const T* __restrict__ box, int yl, int xl, T (*acc)[PxbD3Cfg<T, K>::R][PxbD3Cfg<T, K>::VEC],
                              int a_lo, int a_hi) {
PXB_NOUNROLL
    for (int bp = 0; bp < K; ++bp) pxb_d3_row_some<T, K>(coef + bp * PxbD3Cfg<T, K>::CROW, box, yl, xl, bp, acc, a_lo, a_hi);
}

// the epilogue's `add` samples of output plane q for this thread
template <class T, int K>
PXB_HD void pxb_d3_load_add(const PxbD3P<T, K>& p, T (*addv)[PxbD3Cfg<T, K>::VEC], int64_t b, int q, int y0, int x0, int yl, int xl) {
    using C = PxbD3Cfg<T, K>;
    constexpr int VEC = C::VEC;
    const int x = x0 + xl;
    for (int r = 0; r < C::R; ++r) {
        const int y = y0 + yl + r;
        for (int j = 0; j < VEC; ++j) addv[r][j] = T(0);
        if (!p.add || x >= p.n2 || y >= p.n1) continue;
        const int64_t al = ((b * p.n0 + q) * (int64_t)p.n1 + y) * p.n2 + x;
        if (p.add_period <= 0) {
            const PxbVec<T, VEC> a = pxb_vload<T, VEC>(p.add + al);
            for (int j = 0; j < VEC; ++j) addv[r][j] = a.v[j];
        } else {
            for (int j = 0; j < VEC; ++j) addv[r][j] = p.add[(al + j) % p.add_period];
        }
    }
}

// the finished slot (K - 1) -> output plane q, then every slot moves up by one and slot 0 starts empty
template <class T, int K>
PXB_HD void pxb_d3_emit(const PxbD3P<T, K>& p, T* __restrict__ out, const T (*fin)[PxbD3Cfg<T, K>::VEC], const T (*addv)[PxbD3Cfg<T, K>::VEC], int64_t b,
                        int q, int y0, int x0, int yl, int xl) {
    using C = PxbD3Cfg<T, K>;
    constexpr int VEC = C::VEC;
    const int x = x0 + xl;
    if (x >= p.n2) return;
    const int64_t s0 = (int64_t)p.n1 * p.n2;
    for (int r = 0; r < C::R; ++r) {
        const int y = y0 + yl + r;
        if (y >= p.n1) break;
        PxbVec<T, VEC> o;
        for (int j = 0; j < VEC; ++j) {
            o.v[j] = p.alpha * fin[r][j];
            if (p.add) o.v[j] += p.beta * addv[r][j];
        }
        pxb_vstore<T, VEC>(out + b * p.vol + (int64_t)q * s0 + (int64_t)y * p.n2 + x, o);
    }
}

template <class T, int K>
PXB_HD void pxb_d3_shift(T (*acc)[PxbD3Cfg<T, K>::R][PxbD3Cfg<T, K>::VEC]) {
    using C = PxbD3Cfg<T, K>;
PXB_UNROLL
    for (int a = K - 1; a > 0; --a)
        for (int r = 0; r < C::R; ++r)
            for (int j = 0; j < C::VEC; ++j) acc[a][r][j] = acc[a - 1][r][j];
    for (int r = 0; r < C::R; ++r)
        for (int j = 0; j < C::VEC; ++j) acc[0][r][j] = T(0);
}

// The cube the launcher runs a k0 x k1 x k2 kernel in: 3, 5 or 7 taps per axis, zero taps behind the kernel's own (the centre
// stays where it is).  0 when the kernel does not fit or the zero taps would more than double the work.
inline int pxb_d3_cube(const int* ksize) {
    const int m = ksize[0] > ksize[1] ? (ksize[0] > ksize[2] ? ksize[0] : ksize[2]) : (ksize[1] > ksize[2] ? ksize[1] : ksize[2]);
    const int K = m <= 3 ? 3 : m <= 5 ? 5 : m <= 7 ? 7 : 0;
    if (!K || 2 * ksize[0] * ksize[1] * ksize[2] < K * K * K) return 0;
    return K;
}

// host: chunk length along the marching axis.  A chunk reads K - 1 planes it shares with its neighbours (half of their FMAs
// belong to the neighbour's outputs and are skipped); fewer chunks leave the last wave of CTAs (2 per SM) partly empty.
// Smallest cost of both.
template <class T, int K>
inline int pxb_d3_setup(PxbD3P<T, K>& p) {
    using C = PxbD3Cfg<T, K>;
    if (p.n2 % C::VEC) return 2;
    if (p.c0 < 0 || p.c0 >= K || p.c1 < 0 || p.c1 >= K || p.c2 < 0 || p.c2 >= K) return 3;
    p.ntx = (p.n2 + C::TX - 1) / C::TX;
    p.nty = (p.n1 + C::TY - 1) / C::TY;
    const int64_t tiles = (int64_t)p.ntx * p.nty * p.batch;
    const double slots = 148.0 * 2.0;
    int best = p.n0;
    double best_cost = 1e300;
    for (int nch = 1; nch <= p.n0; nch = nch < 8 ? nch + 1 : nch * 2) {
        const int chunk = (p.n0 + nch - 1) / nch;
        if (chunk < K && nch > 1) break;
        const int64_t ctas = tiles * ((p.n0 + chunk - 1) / chunk);
        const double waves = (double)ctas / slots;
        const double tail = (double)(int64_t)(waves + 0.999999) / waves;
        const double cost = tail * (1.0 + 0.5 * (double)(K - 1) / chunk);  // (the end planes run the slots of the chunk only)
        if (cost < best_cost - 1e-9) { best_cost = cost; best = chunk; }
    }
    p.chunk = best;
    p.nchunk = (p.n0 + best - 1) / best;
    if (tiles * p.nchunk > 0x7fffffffLL) return 5;
    return 0;
}
