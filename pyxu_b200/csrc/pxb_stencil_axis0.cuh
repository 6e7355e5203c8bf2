// pxb_stencil_axis0.cuh -- per-thread body of the streaming axis-0 stencil (design: pxb_stencil_axis0.cu).  __host__ __device__:
// tests/emu replays it column by column.
#pragma once
#include "pxb_tv_fast.cuh"

struct Axis0P {
    int n0;              // owned planes
    int64_t plane;       // elements per plane
    int64_t vol;         // elements between batch items
    int c0, chunk, nchunk;
    int lo_planes, hi_planes;  // readable planes below plane 0 / above plane n0-1 (open slab sides)
    double coef[16];
    // FOLD instances only (folding boundary mode along axis 0, single-domain arrays):
    int mode;            // pxb_mode of axis 0
    int adjoint;         // 0: S o Pad (out-of-domain input planes are the planes the boundary map folds them onto)
                         // 1: Pad^T o S0^T: `coef` / c0 hold the REVERSED taps and the mirrored centre; every output plane also
                         //    collects S0^T at the padded planes that fold onto it
    int pad_lo, pad_hi;  // pad widths of the operator itself (centre, k0 - 1 - centre of the un-reversed kernel)
};

// one thread = one 16-byte column, marching through the planes [m0, m1) of its chunk with the last K0 planes in registers
template <class T, int VEC, int K0, bool FOLD>
PXB_HD void pxb_axis0_column(const Axis0P& p, const T* __restrict__ src, T* __restrict__ dst, int m0, int m1) {
    T c[K0];
    for (int j = 0; j < K0; ++j) c[j] = T(p.coef[j]);
    PxbVec<T, VEC> ring[K0];
    for (int j = 0; j < K0; ++j)
        for (int v = 0; v < VEC; ++v) ring[j].v[v] = T(0);
    const int last = m1 + K0 - 1 - p.c0;  // one past the last input plane
    for (int pl = m0 - p.c0; pl < last; ++pl) {
        PxbVec<T, VEC> t;
        int pe = pl;
        bool have = pl >= -p.lo_planes && pl < p.n0 + p.hi_planes;
        if (FOLD && !have && !p.adjoint && p.mode != PXB_CONSTANT) {  // S o Pad: the padded plane is a copy of plane m(pl)
            pe = pxb_bmap(pl, p.n0, p.mode);
            have = pe != PXB_NOSRC;
        }
        if (have) t = pxb_vload<T, VEC>(src + (int64_t)pe * p.plane);
        else for (int v = 0; v < VEC; ++v) t.v[v] = T(0);
        for (int j = 0; j + 1 < K0; ++j) ring[j] = ring[j + 1];
        ring[K0 - 1] = t;
        const int q = pl - (K0 - 1 - p.c0);
        if (q >= m0) {
            PxbVec<T, VEC> o;
            for (int v = 0; v < VEC; ++v) {
                T a = T(0);
                for (int j = 0; j < K0; ++j) a += c[j] * ring[j].v[v];
                o.v[v] = a;
            }
            if (FOLD && p.adjoint && p.mode != PXB_CONSTANT) {
                // Pad^T: + S0^T y at every padded plane e (outside the array, within the pad widths) with m(e) = q;
                // S0^T y [e] = sum_j c[j] y[e - c0 + j] over the planes inside the array.  Only the planes next to a face get here.
                PxbPre P;
                pxb_preimage(q, p.n0, p.mode, pxb_imax(p.pad_lo, p.pad_hi), 0, 0, P);
                for (int a = 1; a < P.cnt; ++a) {
                    const int e_lo = pxb_imax(P.lo[a], -p.pad_lo), e_hi = pxb_imin(P.hi[a], p.n0 + p.pad_hi - 1);
                    for (int e = e_lo; e <= e_hi; ++e)
                        for (int j = 0; j < K0; ++j) {
                            const int s = e - p.c0 + j;
                            if (s < 0 || s >= p.n0) continue;
                            const PxbVec<T, VEC> y = pxb_vload<T, VEC>(src + (int64_t)s * p.plane);
                            for (int v = 0; v < VEC; ++v) o.v[v] += c[j] * y.v[v];
                        }
                }
            }
            pxb_vstore<T, VEC>(dst + (int64_t)q * p.plane, o);
        }
    }
}
