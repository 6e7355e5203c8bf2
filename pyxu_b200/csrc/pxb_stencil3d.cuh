// pxb_stencil3d.cuh -- separable 3-D stencil ('constant' boundaries) in ONE pass over HBM.
//
// The reference runs a separable kernel as a chain of three 1-D stencils (stencil.py:497-538): 24 B/voxel in fp32.
// Here a CTA owns a tile of TY x TX samples and MARCHES along the slowest axis: per input plane it stages the tile's
// window by TMA (two stages in flight, zero fill = 'constant' Pad), filters it in-plane (row pass shared -> shared,
// column pass shared -> registers) and pushes the result into a ring of the last K0 in-plane-filtered planes that
// lives in REGISTERS (each thread only ever needs its own samples of those planes); the output plane is the K0-tap
// combination of the ring.  8 B/voxel (+ (K0-1)/chunk for the planes two chunks share).
// Slab cuts along the marching axis read the neighbour's ghost planes.  Bodies are __host__ __device__ (tests/emu).
#pragma once
#include "pxb_stencil_tma.cuh"

struct PxbSt3P {
    PxbSt2P s;             // in-plane geometry / factors / epilogue (n1, n2, k1, k2, bw, bh, ntx, nty, coef1, coef2, alpha, beta, add ...)
    int n0;                // owned planes
    int64_t batch, vol;    // batch items, elements between them (plane_alloc * n1 * n2)
    int k0, c0;
    double coef0[PXB_ST2_MAXTAP];
    int lo_planes, hi_planes;  // readable ghost planes below plane 0 / above plane n0-1
    int chunk, nchunk;
};

template <class T, int VEC>
struct PxbSt3Cfg {
    static constexpr int TXL = 32, TX = TXL * VEC, TY = 16, R = 2, NT = 256;
};

// row pass item of the 3-D kernel: same arithmetic as pxb_st2_row_item with the 3-D tile height
template <class T, int VEC, int NV>
PXB_HD void pxb_st3_row_item(const PxbSt3P& p, const T* __restrict__ box, T* __restrict__ mid, int y, int xl, const T* c2) {
    using C = PxbSt3Cfg<T, VEC>;
    T v[NV * VEC];
    const T* __restrict__ src = box + y * p.s.bw + xl;
    for (int n = 0; n < NV; ++n) {
        const PxbVec<T, VEC> t = pxb_vload<T, VEC>(src + n * VEC);
        for (int j = 0; j < VEC; ++j) v[n * VEC + j] = t.v[j];
    }
    PxbVec<T, VEC> acc;
    for (int j = 0; j < VEC; ++j) acc.v[j] = T(0);
    for (int q = 0; q < NV * VEC - VEC + 1; ++q)
        for (int j = 0; j < VEC; ++j) acc.v[j] += c2[q] * v[q + j];
    pxb_vstore<T, VEC>(mid + y * C::TX + xl, acc);
}

// column pass: R adjacent rows x VEC columns of the in-plane-filtered plane (c1p: factor padded with R-1 zeros each side)
template <class T, int VEC>
PXB_HD void pxb_st3_col_item(const PxbSt3P& p, const T* __restrict__ mid, int yl, int xl, const T* __restrict__ c1p, T (*t)[VEC]) {
    using C = PxbSt3Cfg<T, VEC>;
    for (int r = 0; r < C::R; ++r)
        for (int j = 0; j < VEC; ++j) t[r][j] = T(0);
    T cw[C::R];
    for (int r = 0; r < C::R; ++r) cw[r] = T(0);
    const T* __restrict__ src = mid + yl * C::TX + xl;
    for (int i = 0; i < C::R + p.s.k1 - 1; ++i) {
        for (int r = C::R - 1; r > 0; --r) cw[r] = cw[r - 1];
        cw[0] = c1p[C::R - 1 + i];
        const PxbVec<T, VEC> v = pxb_vload<T, VEC>(src + i * C::TX);
        for (int r = 0; r < C::R; ++r)
            for (int j = 0; j < VEC; ++j) t[r][j] += cw[r] * v.v[j];
    }
}

// The ring of the last K0 in-plane-filtered planes is indexed CIRCULARLY with compile-time slots: the marching loop is
// unrolled K0 times, plane u of a group writes slot u, and the K0-tap combination reads slot (u + 1 + k) % K0 for tap k
// (oldest plane first).  (A shifting ring cost K0*R*VEC register moves per plane: 7 of the 64 instructions per voxel.)
// `addv`: the epilogue's `add` samples, loaded by the caller ahead of the passes (their latency was the top stall).
template <class T, int VEC, int K0>
PXB_HD void pxb_st3_store(const PxbSt3P& p, T* __restrict__ out, const T* c0v, const T (*ring)[PxbSt3Cfg<T, VEC>::R][VEC], int u, const T (*addv)[VEC],
                          int64_t b, int q, int y0, int x0, int yl, int xl) {
    using C = PxbSt3Cfg<T, VEC>;
    const int x = x0 + xl;
    if (x >= p.s.n2) return;
    const T alpha = T(p.s.alpha), beta = T(p.s.beta);
    const int64_t s0 = (int64_t)p.s.n1 * p.s.n2;
    for (int r = 0; r < C::R; ++r) {
        const int y = y0 + yl + r;
        if (y >= p.s.n1) break;
        PxbVec<T, VEC> o;
        for (int j = 0; j < VEC; ++j) {
            T a = T(0);
            for (int k = 0; k < K0; ++k) a += c0v[k] * ring[(u + 1 + k) % K0][r][j];
            o.v[j] = alpha * a;
            if (p.s.add) o.v[j] += beta * addv[r][j];
        }
        pxb_vstore<T, VEC>(out + b * p.vol + (int64_t)q * s0 + (int64_t)y * p.s.n2 + x, o);
    }
}

// the epilogue's `add` samples of output plane q for this thread (`add` is dense (batch, n0, n1, n2): no ghost planes)
template <class T, int VEC>
PXB_HD void pxb_st3_load_add(const PxbSt3P& p, T (*addv)[VEC], int64_t b, int q, int y0, int x0, int yl, int xl) {
    using C = PxbSt3Cfg<T, VEC>;
    const T* __restrict__ add = (const T*)p.s.add;
    const int x = x0 + xl;
    for (int r = 0; r < C::R; ++r) {
        const int y = y0 + yl + r;
        for (int j = 0; j < VEC; ++j) addv[r][j] = T(0);
        if (!add || x >= p.s.n2 || y >= p.s.n1) continue;
        const int64_t al = ((b * p.n0 + q) * (int64_t)p.s.n1 + y) * p.s.n2 + x;
        if (p.s.add_period <= 0) {
            const PxbVec<T, VEC> a = pxb_vload<T, VEC>(add + al);
            for (int j = 0; j < VEC; ++j) addv[r][j] = a.v[j];
        } else {
            for (int j = 0; j < VEC; ++j) addv[r][j] = add[(al + j) % p.s.add_period];
        }
    }
}

// host: geometry.  p.s must hold n1, n2, k1, k2, c1, c2, coef1, coef2 (unpadded); fills the rest.  0 or a reason code.
template <class T, int VEC>
inline int pxb_st3_setup(PxbSt3P& p) {
    using C = PxbSt3Cfg<T, VEC>;
    p.s.dense = 0;
    p.s.nimg = 1;
    if (int why = pxb_st2_setup<T, VEC>(p.s)) return why;   // pads the column factor, sets bw / ntx ... for TY = 32
    p.s.bh = C::TY + p.s.k1 - 1;                            // ... the 3-D tile is 16 rows high
    p.s.nty = (p.s.n1 + C::TY - 1) / C::TY;
    if (p.k0 != 3 && p.k0 != 5 && p.k0 != 7 && p.k0 != 9) return 11;  // compiled ring depths
    if (p.c0 < 0 || p.c0 >= p.k0) return 3;
    int chunk = p.n0 < 64 ? p.n0 : 64;
    const int64_t tiles = (int64_t)p.s.ntx * p.s.nty * p.batch;
    while (chunk > 16 && tiles * ((p.n0 + chunk - 1) / chunk) < 148 * 8) chunk = (chunk + 1) / 2;
    p.chunk = chunk;
    p.nchunk = (p.n0 + chunk - 1) / chunk;
    if (tiles * p.nchunk > 0x7fffffffLL) return 5;
    return 0;
}
