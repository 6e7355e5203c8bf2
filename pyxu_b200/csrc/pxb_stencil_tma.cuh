// pxb_stencil_tma.cuh -- shared-memory-tiled 2-D stencil ('constant' boundaries) with TMA-staged halos.
//
// One CTA = one tile of TY x TX outputs of one image.  A single TMA box load brings the (TY + k1 - 1) x (TX + k2 - 1)
// input window into shared memory; samples outside the image are ZERO-FILLED by the TMA unit, which is the reference's
// 'constant' Pad (pad.py:252-258), so no thread ever tests a boundary.  Then either
//   * separable (k = k1 (x) k2, what the reference runs as a CHAIN of 1-D stencils, one HBM round trip each:
//     stencil.py:497-538): row pass shared -> shared, column pass shared -> registers: ONE HBM round trip;
//   * dense k1 x k2: register-blocked accumulation (4 rows x VEC columns per thread) out of shared memory.
// Epilogue: out = alpha * S(in) + beta * add[i % period]  (fuses the "- y" of a data term A x - y).
// HBM traffic: 8 B/voxel (fp32) + the halo re-reads, which are L2 hits (neighbouring tiles run together).
//
// The per-thread bodies are __host__ __device__; tests/emu replays them with the box load emulated.
#pragma once
#include "pxb_tv_fast.cuh"

#define PXB_ST2_MAXTAP 16

struct PxbSt2P {          // by-value kernel parameter
    int n1, n2;           // image rows, columns (the OUTPUT grid)
    int64_t nimg;
    int k1, k2, c1, c2;   // taps / centers along rows (axis 1) and columns (axis 2) -- k2 / c2 AFTER the alignment padding below;
                          // the window of output tile (y0, x0) starts at input sample (y0 - c1, x0 - c2): with an input of another
                          // extent the origin of the output grid is folded into c1 / c2, which then may leave [0, k)
    int k2src, extra;     // TMA box starts must be 16-byte aligned in global memory (measured: a start at c0 - 3 floats
                          // faults): the column factor gets `extra` leading zero taps so that its center is a multiple of VEC
    int dense;
    int bw, bh;           // input box: columns (multiple of VEC), rows = TY + k1 - 1
    int ntx, nty;         // tiles per image
    int tpc, ngx;         // consecutive tiles along x one CTA works through (double-buffered box loads), CTAs per tile row
    double coef1[PXB_ST2_MAXTAP], coef2[PXB_ST2_MAXTAP];
    const void* coef;     // dense: device pointer to k1*k2 coefficients
    double alpha, beta;
    const void* add;
    int64_t add_period;
    // prologue (PRO): the stencil acts on  pa*in + pb*in2  (FISTA extrapolation y = (1+a) x - a x_prev), formed in shared memory
    double pa, pb;
    // epilogue kind 1 (proximal-gradient step):  v = ea*e1[i] + eb*e2[i] + alpha*S(.)[i];  out = prox_{tau g}(v);
    //   norms[2*row] += (out - e1)^2, norms[2*row+1] += e1^2   (row = img / imgs_per_row)      -- pgd.py:179-191, stop.py:353-382
    int epi;
    const void* e1;
    const void* e2;
    double ea, eb;
    int gkind;
    double gp0, gp1, tau;
    double* norms;
    int64_t imgs_per_row;
};

// Host side only: extent of the input images when it differs from the output grid (pxb_stencil2d::in_shape).  It shapes
// the tensor map, whose zero fill is the 'constant' extension of the input; the kernels never see it.
struct PxbSt2In {
    int n1, n2;
};

template <class T, int VEC>
struct PxbSt2Cfg {
    static constexpr int TXL = 32, TX = TXL * VEC, TY = 32, R = 4, NT = TXL * (TY / R);
};

// row pass for one vector item: t[y][x..x+VEC) = sum_q c2[q] * in[y][x + q ...]; NV vectors cover the window
// (coefficients arrive as T: converting the by-value doubles inside the tap loops cost an F2F per tap -- the first
// version issued 67 instructions per voxel and was issue-bound at 82 %)
template <class T, int VEC, int NV>
PXB_HD void pxb_st2_row_item(const PxbSt2P& p, const T* __restrict__ box, T* __restrict__ mid, int y, int xl, const T* c2) {
    using C = PxbSt2Cfg<T, VEC>;
    T v[NV * VEC];
    const T* __restrict__ src = box + y * p.bw + xl;
    for (int n = 0; n < NV; ++n) {
        const PxbVec<T, VEC> t = pxb_vload<T, VEC>(src + n * VEC);
        for (int j = 0; j < VEC; ++j) v[n * VEC + j] = t.v[j];
    }
    PxbVec<T, VEC> acc;
    for (int j = 0; j < VEC; ++j) acc.v[j] = T(0);
    // every tap the window covers, unconditionally: c2[q] == 0 beyond the kernel (<= VEC-1 wasted taps, no predicates)
    for (int q = 0; q < NV * VEC - VEC + 1; ++q)
        for (int j = 0; j < VEC; ++j) acc.v[j] += c2[q] * v[q + j];
    pxb_vstore<T, VEC>(mid + y * C::TX + xl, acc);
}

// column pass for one thread: R adjacent rows x VEC columns.  `c1p` = the row factor padded with R-1 zeros on each
// side (c1p[R-1+q] = c1[q]), so that the R running outputs take every loaded row without a range test:
//   out[r] += c1[i - r] * t[i]   for all r   (the coefficient window slides through registers)
template <class T, int VEC>
PXB_HD void pxb_st2_col_item(const PxbSt2P& p, const T* __restrict__ mid, int yl, int xl, const T* __restrict__ c1p, T (*acc)[VEC]) {
    using C = PxbSt2Cfg<T, VEC>;
    for (int r = 0; r < C::R; ++r)
        for (int j = 0; j < VEC; ++j) acc[r][j] = T(0);
    T cw[C::R];  // cw[r] = c1[i - r]
    for (int r = 0; r < C::R; ++r) cw[r] = T(0);
    const T* __restrict__ src = mid + yl * C::TX + xl;
    for (int i = 0; i < C::R + p.k1 - 1; ++i) {
        for (int r = C::R - 1; r > 0; --r) cw[r] = cw[r - 1];
        cw[0] = c1p[C::R - 1 + i];
        const PxbVec<T, VEC> t = pxb_vload<T, VEC>(src + i * C::TX);
        for (int r = 0; r < C::R; ++r)
            for (int j = 0; j < VEC; ++j) acc[r][j] += cw[r] * t.v[j];
    }
}

// dense k1 x k2 for one thread; `ck`: the coefficients in shared memory (row-major k1 x k2)
template <class T, int VEC, int NV>
PXB_HD void pxb_st2_dense_item(const PxbSt2P& p, const T* __restrict__ box, const T* __restrict__ ck, int yl, int xl, T (*acc)[VEC]) {
    using C = PxbSt2Cfg<T, VEC>;
    for (int r = 0; r < C::R; ++r)
        for (int j = 0; j < VEC; ++j) acc[r][j] = T(0);
    for (int q1 = 0; q1 < p.k1; ++q1) {
        T c[NV * VEC - VEC + 1];
        for (int q = 0; q < NV * VEC - VEC + 1; ++q) c[q] = q < p.k2 ? ck[q1 * p.k2 + q] : T(0);
        for (int r = 0; r < C::R; ++r) {
            T v[NV * VEC];
            const T* __restrict__ src = box + (yl + r + q1) * p.bw + xl;
            for (int n = 0; n < NV; ++n) {
                const PxbVec<T, VEC> t = pxb_vload<T, VEC>(src + n * VEC);
                for (int j = 0; j < VEC; ++j) v[n * VEC + j] = t.v[j];
            }
            for (int q = 0; q < NV * VEC - VEC + 1; ++q)
                for (int j = 0; j < VEC; ++j) acc[r][j] += c[q] * v[q + j];
        }
    }
}

// The epilogue's global operands (`add`; x / x_prev of the prox step) are loaded by pxb_st2_load_epi BEFORE the thread
// waits for its TMA box, so their DRAM latency overlaps the box load and the passes (ncu on the first version: the
// epilogue loads were the top stall, long-scoreboard, of every instance with an epilogue operand).
template <class T, int VEC>
struct PxbSt2Epi {
    T a[PxbSt2Cfg<T, VEC>::R][VEC];  // add            | x      (prox step)
    T b[PxbSt2Cfg<T, VEC>::R][VEC];  //                | x_prev (prox step)
};

// L2 prefetch of the same operands, for the dense (register-hungry, FMA-bound) instances: holding the operands in
// registers across the accumulation cost them occupancy (FISTA 5x5: 1.76 -> 1.91 ms), a prefetch costs nothing.
PXB_HD void pxb_prefetch_l2(const void* ptr) {
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
#else
    (void)ptr;
#endif
}
template <class T, int VEC>
PXB_HD void pxb_st2_prefetch_epi(const PxbSt2P& p, int64_t img, int y0, int x0, int yl, int xl) {
    using C = PxbSt2Cfg<T, VEC>;
    const int x = x0 + xl;
    const T* __restrict__ pa = (const T*)(p.epi == 1 ? p.e1 : p.add);
    const T* __restrict__ pb = (const T*)(p.epi == 1 ? p.e2 : nullptr);
    if (x >= p.n2 || (!pa && !pb) || (p.epi != 1 && p.add_period > 0)) return;
    for (int r = 0; r < C::R; ++r) {
        const int y = y0 + yl + r;
        if (y >= p.n1) break;
        const int64_t lin = (img * p.n1 + y) * (int64_t)p.n2 + x;
        if (pa) pxb_prefetch_l2(pa + lin);
        if (pb) pxb_prefetch_l2(pb + lin);
    }
}

template <class T, int VEC>
PXB_HD void pxb_st2_load_epi(const PxbSt2P& p, PxbSt2Epi<T, VEC>& e, int64_t img, int y0, int x0, int yl, int xl) {
    using C = PxbSt2Cfg<T, VEC>;
    const int x = x0 + xl;
    const T* __restrict__ pa = (const T*)(p.epi == 1 ? p.e1 : p.add);
    const T* __restrict__ pb = (const T*)(p.epi == 1 ? p.e2 : nullptr);
    for (int r = 0; r < C::R; ++r) {
        for (int j = 0; j < VEC; ++j) e.a[r][j] = e.b[r][j] = T(0);
        const int y = y0 + yl + r;
        if (x >= p.n2 || y >= p.n1) continue;
        const int64_t lin = (img * p.n1 + y) * (int64_t)p.n2 + x;
        if (pa) {
            if (p.epi == 1 || p.add_period <= 0) {
                const PxbVec<T, VEC> v = pxb_vload<T, VEC>(pa + lin);
                for (int j = 0; j < VEC; ++j) e.a[r][j] = v.v[j];
            } else {
                for (int j = 0; j < VEC; ++j) e.a[r][j] = pa[(lin + j) % p.add_period];
            }
        }
        if (pb) {
            const PxbVec<T, VEC> v = pxb_vload<T, VEC>(pb + lin);
            for (int j = 0; j < VEC; ++j) e.b[r][j] = v.v[j];
        }
    }
}

// epilogue + store of one thread's R x VEC outputs (tile origin (y0, x0) of image `img`)
template <class T, int VEC>
PXB_HD void pxb_st2_store(const PxbSt2P& p, T* __restrict__ out, const PxbSt2Epi<T, VEC>& e, int64_t img, int y0, int x0, int yl, int xl, T (*acc)[VEC]) {
    using C = PxbSt2Cfg<T, VEC>;
    const int x = x0 + xl;
    if (x >= p.n2) return;
    const T alpha = T(p.alpha), beta = T(p.beta);
    for (int r = 0; r < C::R; ++r) {
        const int y = y0 + yl + r;
        if (y >= p.n1) break;
        const int64_t lin = (img * p.n1 + y) * (int64_t)p.n2 + x;
        PxbVec<T, VEC> o;
        for (int j = 0; j < VEC; ++j) o.v[j] = alpha * acc[r][j];
        if (p.add)
            for (int j = 0; j < VEC; ++j) o.v[j] += beta * e.a[r][j];
        pxb_vstore<T, VEC>(out + lin, o);
    }
}

// proximal-gradient epilogue for one thread's R x VEC outputs; returns the RelError partial sums in nrm[0..1]
template <class T, int VEC>
PXB_HD void pxb_st2_store_prox(const PxbSt2P& p, T* __restrict__ out, const PxbSt2Epi<T, VEC>& e, int64_t img, int y0, int x0, int yl, int xl,
                               T (*acc)[VEC], double* nrm) {
    using C = PxbSt2Cfg<T, VEC>;
    const int x = x0 + xl;
    if (x >= p.n2) return;
    const T alpha = T(p.alpha), ea = T(p.ea), eb = T(p.eb), tau = T(p.tau), gp0 = T(p.gp0), gp1 = T(p.gp1);
    for (int r = 0; r < C::R; ++r) {
        const int y = y0 + yl + r;
        if (y >= p.n1) break;
        const int64_t lin = (img * p.n1 + y) * (int64_t)p.n2 + x;
        PxbVec<T, VEC> o;
        for (int j = 0; j < VEC; ++j) {
            const T v = ea * e.a[r][j] + eb * e.b[r][j] + alpha * acc[r][j];
            o.v[j] = pxb_prox_eval<T>(p.gkind, gp0, gp1, v, tau);
            if (p.norms) {
                const double dd = (double)o.v[j] - (double)e.a[r][j];
                nrm[0] += dd * dd;
                nrm[1] += (double)e.a[r][j] * (double)e.a[r][j];
            }
        }
        pxb_vstore<T, VEC>(out + lin, o);  // (`out` may be the x_prev buffer: its samples were read into e.b above)
    }
}

// dense coefficients as the kernel sees them (row-major k1 x k2 with the leading zero taps): element i
template <class T>
PXB_HD T pxb_st2_dense_coef(const PxbSt2P& p, const T* __restrict__ ck, int i) {
    const int q1 = i / p.k2, q = i - q1 * p.k2;
    return q >= p.extra ? ck[q1 * p.k2src + q - p.extra] : T(0);
}

// prologue: window = pa*box + pb*box2, vector item `it` of the (bh x bw) window
template <class T, int VEC>
PXB_HD void pxb_st2_combine_item(const PxbSt2P& p, T* __restrict__ box, const T* __restrict__ box2, int it) {
    const PxbVec<T, VEC> a = pxb_vload<T, VEC>(box + it * VEC), b = pxb_vload<T, VEC>(box2 + it * VEC);
    PxbVec<T, VEC> o;
    for (int j = 0; j < VEC; ++j) o.v[j] = T(p.pa) * a.v[j] + T(p.pb) * b.v[j];
    pxb_vstore<T, VEC>(box + it * VEC, o);
}

// number of VEC-wide vectors that cover a window of k taps starting at a vector boundary
PXB_HD int pxb_st2_nv(int k, int vec) { return (k - 1 + vec + vec - 1) / vec; }

// host: geometry.  Returns 0 or a reason code when outside the envelope.
template <class T, int VEC>
inline int pxb_st2_setup(PxbSt2P& p, const PxbSt2In* in = nullptr) {
    using C = PxbSt2Cfg<T, VEC>;
    if (p.k1 < 1 || p.k2 < 1 || p.k1 > PXB_ST2_MAXTAP || p.k2 > PXB_ST2_MAXTAP) return 1;
    // (with an input of another extent the centers were validated before the origin was folded into them)
    if (!in && (p.c1 < 0 || p.c1 >= p.k1 || p.c2 < 0 || p.c2 >= p.k2)) return 3;
    if (in && in->n2 % VEC) return 2;
    p.k2src = p.k2;
    p.extra = (VEC - p.c2 % VEC) % VEC;
    if (p.k2 + p.extra > PXB_ST2_MAXTAP) return 1;
    for (int q = p.k2 - 1; q >= 0; --q) p.coef2[q + p.extra] = p.coef2[q];
    for (int q = 0; q < p.extra; ++q) p.coef2[q] = 0.0;
    p.k2 += p.extra;
    p.c2 += p.extra;
    for (int q = p.k2; q < PXB_ST2_MAXTAP; ++q) p.coef2[q] = 0.0;  // the row pass runs every tap its window covers
    for (int q = p.k1; q < PXB_ST2_MAXTAP; ++q) p.coef1[q] = 0.0;
    if (pxb_st2_nv(p.k2, VEC) > (VEC == 4 ? 4 : 6)) return 1;  // compiled window widths: 13 taps (fp32), 11 taps (fp64)
    if (p.n2 % VEC) return 2;
    p.bh = C::TY + p.k1 - 1;
    p.bw = C::TX + (pxb_st2_nv(p.k2, VEC) - 1) * VEC;
    if (p.bw > 256 || p.bh > 256) return 4;
    p.ntx = (p.n2 + C::TX - 1) / C::TX;
    p.nty = (p.n1 + C::TY - 1) / C::TY;
    // separable kernels are load-latency bound: 4 tiles per CTA with the next box in flight (measured 8192^2 9x9: 0.132 -> 0.123 ms);
    // dense kernels are FMA-bound and lose occupancy and wave balance to the second stage (0.264 -> 0.310 ms): one tile per CTA
    p.tpc = p.dense ? 1 : (p.ntx < 4 ? p.ntx : 4);
    p.ngx = (p.ntx + p.tpc - 1) / p.tpc;
    if ((int64_t)p.ngx * p.nty * p.nimg > 0x7fffffffLL) return 5;
    return 0;
}
