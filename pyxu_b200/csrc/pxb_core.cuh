// pxb_core.cuh -- per-sample arithmetic of the hot path, shared by every kernel.
//
// Everything here is `__host__ __device__` on purpose: the CUDA kernels in pxb_kernels.cu call
// these functions once per output sample, and tests/emu compiles the very same functions with g++
// to check index handling against the oracle on machines without a GPU (test infrastructure only;
// the shipped library has no host compute path).
#pragma once
#include <stdint.h>
#include <math.h>

#include "../../include/pyxu_b200.h"

#if defined(__CUDACC__)
#define PXB_HD __host__ __device__ __forceinline__
#else
#define PXB_HD inline
#endif

// ---------------------------------------------------------------------------------------------
// Boundary handling.
// Reference semantics: Stencil = Trim o S0 o Pad (src/pyxu/operator/linop/stencil/stencil.py:76-84),
// with Pad the numpy.pad modes (src/pyxu/operator/linop/pad.py:252-302).  Pad widths are bounded
// (pad.py:217-229: wrap<=N, reflect<=N-1, symmetric<=N) so a coordinate folds at most once.
// pxb_bmap maps an extended coordinate e in [-p, n+p) to the source sample, or -1 (contributes 0).
// `open_lo/open_hi`: that side is not a domain boundary (slab interior): halo planes are addressed
// directly with their out-of-range coordinate.
// ---------------------------------------------------------------------------------------------
#define PXB_NOSRC (-(1 << 30))

PXB_HD int pxb_bmap(int e, int n, int mode, int open_lo = 0, int open_hi = 0) {
    if ((unsigned)e < (unsigned)n) return e;
    if (e < 0) {
        if (open_lo) return e;
        switch (mode) {
            case PXB_CONSTANT: return PXB_NOSRC;
            case PXB_WRAP: return e + n;
            case PXB_REFLECT: return -e;
            case PXB_SYMMETRIC: return -e - 1;
            default: return 0;
        }
    } else {
        if (open_hi) return e;
        switch (mode) {
            case PXB_CONSTANT: return PXB_NOSRC;
            case PXB_WRAP: return e - n;
            case PXB_REFLECT: return 2 * (n - 1) - e;
            case PXB_SYMMETRIC: return 2 * n - 1 - e;
            default: return n - 1;
        }
    }
}

// Pre-image of sample s under the boundary map: all extended coordinates e in [-p, n+p) with
// bmap(e) == s.  This is the transpose of Pad (pad.py:307-375): the adjoint of a stencil gathers
// through these.  Up to 3 intervals [lo, hi]; interval 0 is always {s}.
struct PxbPre {
    int lo[3], hi[3], cnt;
};

PXB_HD void pxb_preimage(int s, int n, int mode, int p, int open_lo, int open_hi, PxbPre& P) {
    P.cnt = 1;
    P.lo[0] = P.hi[0] = s;
    if (mode == PXB_CONSTANT || p <= 0) return;
    int c = 1;
    if (!open_lo) {  // samples folded in from e < 0
        if (mode == PXB_WRAP) {
            if (s >= n - p) { P.lo[c] = P.hi[c] = s - n; ++c; }
        } else if (mode == PXB_REFLECT) {
            if (s >= 1 && s <= p) { P.lo[c] = P.hi[c] = -s; ++c; }
        } else if (mode == PXB_SYMMETRIC) {
            if (s <= p - 1) { P.lo[c] = P.hi[c] = -1 - s; ++c; }
        } else {  // edge
            if (s == 0) { P.lo[c] = -p; P.hi[c] = -1; ++c; }
        }
    }
    if (!open_hi) {  // samples folded in from e >= n
        if (mode == PXB_WRAP) {
            if (s < p) { P.lo[c] = P.hi[c] = s + n; ++c; }
        } else if (mode == PXB_REFLECT) {
            if (s <= n - 2 && s >= n - 1 - p) { P.lo[c] = P.hi[c] = 2 * (n - 1) - s; ++c; }
        } else if (mode == PXB_SYMMETRIC) {
            if (s >= n - p) { P.lo[c] = P.hi[c] = 2 * n - 1 - s; ++c; }
        } else {
            if (s == n - 1) { P.lo[c] = n; P.hi[c] = n - 1 + p; ++c; }
        }
    }
    P.cnt = c;
}

// true when sample s has no folded pre-image (conservative for every mode): only e == s maps onto it.
PXB_HD bool pxb_no_fold(int s, int n, int mode, int p, int open_lo, int open_hi) {
    if (mode == PXB_CONSTANT || p <= 0) return true;
    return (open_lo || s > p) && (open_hi || s < n - 1 - p);
}

PXB_HD int pxb_imax(int a, int b) { return a > b ? a : b; }
PXB_HD int pxb_imin(int a, int b) { return a < b ? a : b; }

// ---------------------------------------------------------------------------------------------
// 1-D correlation along one axis of a strided line (Gradient directions, separable Stencil passes).
//   apply  : y[i] = sum_q c[q] * x[bmap(i - cen + q)]                    (_stencil.py:278-305)
//   adjoint: x[s] = sum_{e in pre(s)} sum_q c[q] * y[e + cen - q],  0 <= e+cen-q < n (or halo)
// `line` points at coordinate 0 of the line; stride in elements.
// ---------------------------------------------------------------------------------------------
template <class T, class C>
PXB_HD T pxb_corr1d_at(const T* __restrict__ line, int64_t stride, int n, int mode, int nt, int cen,
                       const C* __restrict__ coef, int i, int open_lo = 0, int open_hi = 0) {
    T acc = T(0);
    const int e0 = i - cen;
    if (e0 >= 0 && e0 + nt <= n) {
        for (int q = 0; q < nt; ++q) acc += T(coef[q]) * line[(int64_t)(e0 + q) * stride];
    } else {
        for (int q = 0; q < nt; ++q) {
            const int j = pxb_bmap(e0 + q, n, mode, open_lo, open_hi);
            if (j != PXB_NOSRC) acc += T(coef[q]) * line[(int64_t)j * stride];
        }
    }
    return acc;
}

template <class T, class C>
PXB_HD T pxb_corr1d_adj_at(const T* __restrict__ line, int64_t stride, int n, int mode, int nt, int cen,
                           const C* __restrict__ coef, int s, int open_lo = 0, int open_hi = 0) {
    T acc = T(0);
    const int p = nt - 1;
    // rows i that exist: [ilo, ihi]; halo rows count when the side is open.
    const int ilo = open_lo ? -p : 0, ihi = open_hi ? n - 1 + p : n - 1;
    if (s - (nt - 1 - cen) >= 0 && s + cen <= n - 1 && pxb_no_fold(s, n, mode, p, open_lo, open_hi)) {  // interior: single pre-image, all rows valid
        for (int q = 0; q < nt; ++q) acc += T(coef[q]) * line[(int64_t)(s + cen - q) * stride];
        return acc;
    }
    PxbPre P;
    pxb_preimage(s, n, mode, p, open_lo, open_hi, P);
    for (int t = 0; t < P.cnt; ++t) {
        for (int e = P.lo[t]; e <= P.hi[t]; ++e) {
            const int qlo = pxb_imax(0, e + cen - ihi), qhi = pxb_imin(nt - 1, e + cen - ilo);
            for (int q = qlo; q <= qhi; ++q) acc += T(coef[q]) * line[(int64_t)(e + cen - q) * stride];
        }
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------
// Dense N-D stencil (N <= 3) at one output sample.
// ---------------------------------------------------------------------------------------------
struct PxbGeom {  // geometry shared by all kernels: (batch, n0, n1, n2), C-order
    int n0, n1, n2;
    int64_t s0, s1;  // strides of axes 0 and 1 (axis 2 has stride 1); sb = n0*s0 is the batch stride
    int64_t sb;
};

PXB_HD PxbGeom pxb_geom(const int64_t shape[3]) {
    PxbGeom g;
    g.n0 = (int)shape[0]; g.n1 = (int)shape[1]; g.n2 = (int)shape[2];
    g.s1 = shape[2];
    g.s0 = shape[1] * shape[2];
    g.sb = shape[0] * g.s0;
    return g;
}

template <class T>
PXB_HD T pxb_stencil_at(const pxb_stencil_desc& d, const PxbGeom& g, const T* __restrict__ img /*batch item*/,
                        const T* __restrict__ coef, int i0, int i1, int i2) {
    T acc = T(0);
    const int k0 = d.ksize[0], k1 = d.ksize[1], k2 = d.ksize[2];
    const int e0 = i0 - d.center[0], e1 = i1 - d.center[1], e2 = i2 - d.center[2];
    const bool interior = e0 >= 0 && e0 + k0 <= g.n0 && e1 >= 0 && e1 + k1 <= g.n1 && e2 >= 0 && e2 + k2 <= g.n2;
    if (interior) {
        for (int q0 = 0; q0 < k0; ++q0)
            for (int q1 = 0; q1 < k1; ++q1) {
                const T* __restrict__ row = img + (int64_t)(e0 + q0) * g.s0 + (int64_t)(e1 + q1) * g.s1 + e2;
                const T* __restrict__ cf = coef + ((int64_t)q0 * k1 + q1) * k2;
                for (int q2 = 0; q2 < k2; ++q2) acc += cf[q2] * row[q2];
            }
        return acc;
    }
    for (int q0 = 0; q0 < k0; ++q0) {
        const int j0 = pxb_bmap(e0 + q0, g.n0, d.mode[0], d.slab.open_lo, d.slab.open_hi);
        if (j0 == PXB_NOSRC) continue;
        for (int q1 = 0; q1 < k1; ++q1) {
            const int j1 = pxb_bmap(e1 + q1, g.n1, d.mode[1]);
            if (j1 == PXB_NOSRC) continue;
            const T* __restrict__ row = img + (int64_t)j0 * g.s0 + (int64_t)j1 * g.s1;
            const T* __restrict__ cf = coef + ((int64_t)q0 * k1 + q1) * k2;
            for (int q2 = 0; q2 < k2; ++q2) {
                const int j2 = pxb_bmap(e2 + q2, g.n2, d.mode[2]);
                if (j2 != PXB_NOSRC) acc += cf[q2] * row[j2];
            }
        }
    }
    return acc;
}

// Transpose of the above: x[s] = sum_{e in pre(s)} sum_q k[q] * y[e + c - q]   (stencil.py:452-461:
// zero-pad (Trim^T), correlate with the flipped kernel, fold (Pad^T)).
template <class T>
PXB_HD T pxb_stencil_adj_at(const pxb_stencil_desc& d, const PxbGeom& g, const T* __restrict__ img,
                            const T* __restrict__ coef, int s0, int s1, int s2) {
    T acc = T(0);
    const int k0 = d.ksize[0], k1 = d.ksize[1], k2 = d.ksize[2];
    const int c0 = d.center[0], c1 = d.center[1], c2 = d.center[2];
    const bool interior = s0 - (k0 - 1 - c0) >= 0 && s0 + c0 <= g.n0 - 1 && s1 - (k1 - 1 - c1) >= 0 &&
                          s1 + c1 <= g.n1 - 1 && s2 - (k2 - 1 - c2) >= 0 && s2 + c2 <= g.n2 - 1 &&
                          pxb_no_fold(s0, g.n0, d.mode[0], k0 - 1, d.slab.open_lo, d.slab.open_hi) &&
                          pxb_no_fold(s1, g.n1, d.mode[1], k1 - 1, 0, 0) && pxb_no_fold(s2, g.n2, d.mode[2], k2 - 1, 0, 0);
    if (interior) {
        for (int q0 = 0; q0 < k0; ++q0)
            for (int q1 = 0; q1 < k1; ++q1) {
                const T* __restrict__ row = img + (int64_t)(s0 + c0 - q0) * g.s0 + (int64_t)(s1 + c1 - q1) * g.s1 + (s2 + c2);
                const T* __restrict__ cf = coef + ((int64_t)q0 * k1 + q1) * k2;
                for (int q2 = 0; q2 < k2; ++q2) acc += cf[q2] * row[-q2];
            }
        return acc;
    }
    PxbPre P0, P1, P2;
    pxb_preimage(s0, g.n0, d.mode[0], k0 - 1, d.slab.open_lo, d.slab.open_hi, P0);
    pxb_preimage(s1, g.n1, d.mode[1], k1 - 1, 0, 0, P1);
    pxb_preimage(s2, g.n2, d.mode[2], k2 - 1, 0, 0, P2);
    const int i0lo = d.slab.open_lo ? -(k0 - 1) : 0, i0hi = d.slab.open_hi ? g.n0 - 1 + (k0 - 1) : g.n0 - 1;
    for (int t0 = 0; t0 < P0.cnt; ++t0)
        for (int e0 = P0.lo[t0]; e0 <= P0.hi[t0]; ++e0) {
            const int q0lo = pxb_imax(0, e0 + c0 - i0hi), q0hi = pxb_imin(k0 - 1, e0 + c0 - i0lo);
            for (int t1 = 0; t1 < P1.cnt; ++t1)
                for (int e1 = P1.lo[t1]; e1 <= P1.hi[t1]; ++e1) {
                    const int q1lo = pxb_imax(0, e1 + c1 - (g.n1 - 1)), q1hi = pxb_imin(k1 - 1, e1 + c1);
                    for (int t2 = 0; t2 < P2.cnt; ++t2)
                        for (int e2 = P2.lo[t2]; e2 <= P2.hi[t2]; ++e2) {
                            const int q2lo = pxb_imax(0, e2 + c2 - (g.n2 - 1)), q2hi = pxb_imin(k2 - 1, e2 + c2);
                            for (int q0 = q0lo; q0 <= q0hi; ++q0)
                                for (int q1 = q1lo; q1 <= q1hi; ++q1) {
                                    const T* __restrict__ row = img + (int64_t)(e0 + c0 - q0) * g.s0 + (int64_t)(e1 + c1 - q1) * g.s1 + (e2 + c2);
                                    const T* __restrict__ cf = coef + ((int64_t)q0 * k1 + q1) * k2;
                                    for (int q2 = q2lo; q2 <= q2hi; ++q2) acc += cf[q2] * row[-q2];
                                }
                        }
                }
        }
    return acc;
}

// ---------------------------------------------------------------------------------------------
// Gradient stack helpers: direction k of K at one voxel, and (K^T z) at one voxel.
// z layout per batch item: (ndir, n0, n1, n2).
// ---------------------------------------------------------------------------------------------
PXB_HD int64_t pxb_axis_stride(const PxbGeom& g, int axis) { return axis == 0 ? g.s0 : (axis == 1 ? g.s1 : 1); }
PXB_HD int pxb_axis_len(const PxbGeom& g, int axis) { return axis == 0 ? g.n0 : (axis == 1 ? g.n1 : g.n2); }

template <class T>
PXB_HD T pxb_grad_dir_at(const pxb_grad_desc& d, const PxbGeom& g, int k, const T* __restrict__ img, int i0, int i1, int i2) {
    const int ax = d.axis[k];
    const int64_t st = pxb_axis_stride(g, ax);
    const int n = pxb_axis_len(g, ax);
    const int i = ax == 0 ? i0 : (ax == 1 ? i1 : i2);
    const int64_t off = (int64_t)i0 * g.s0 + (int64_t)i1 * g.s1 + i2 - (int64_t)i * st;  // coordinate 0 of the line
    const int ol = ax == 0 ? d.slab.open_lo : 0, oh = ax == 0 ? d.slab.open_hi : 0;
    return pxb_corr1d_at<T, double>(img + off, st, n, d.mode[ax], d.ntap[k], d.center[k], d.coef[k], i, ol, oh);
}

template <class T>
PXB_HD T pxb_grad_adj_at(const pxb_grad_desc& d, const PxbGeom& g, const T* __restrict__ zimg /*(ndir, vol)*/,
                         int64_t comp_stride, int i0, int i1, int i2) {
    T acc = T(0);
    for (int k = 0; k < d.ndir; ++k) {
        const int ax = d.axis[k];
        const int64_t st = pxb_axis_stride(g, ax);
        const int n = pxb_axis_len(g, ax);
        const int i = ax == 0 ? i0 : (ax == 1 ? i1 : i2);
        const int64_t off = (int64_t)i0 * g.s0 + (int64_t)i1 * g.s1 + i2 - (int64_t)i * st;
        const int ol = ax == 0 ? d.slab.open_lo : 0, oh = ax == 0 ? d.slab.open_hi : 0;
        acc += pxb_corr1d_adj_at<T, double>(zimg + k * comp_stride + off, st, n, d.mode[ax], d.ntap[k], d.center[k], d.coef[k], i, ol, oh);
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------
// Proximal maps.
// ---------------------------------------------------------------------------------------------
template <class T>
PXB_HD T pxb_prox_eval(int kind, T p0, T p1, T v, T tau) {
    switch (kind) {
        case PXB_PROX_POS: return v > T(0) ? v : T(0);                      // indicator.py:203-206
        case PXB_PROX_BOX: return v < p0 ? p0 : (v > p1 ? p1 : v);          // indicator.py:58-69 (ord=inf), generalised
        case PXB_PROX_L1: {                                                 // norm.py:47-52 (+ ScaleRule)
            const T a = fabs(v) - p0 * tau;
            return a > T(0) ? (v < T(0) ? -a : a) : T(0);
        }
        case PXB_PROX_POSL1: {                                              // norm.py:400-403
            const T a = v - p0 * tau;
            return a > T(0) ? a : T(0);
        }
        case PXB_PROX_SQL2: return v / (T(2) * (p0 * tau) + T(1));          // norm.py:100-104
        default: return v;
    }
}

// prox_{sigma h*}(p) for one l2-group p[0..G), h = lam*L21 | lam*L1.
// Reference: Moreau identity (abc/operator.py:940-944)  p - sigma * prox_{h/sigma}(p / sigma)  around
// L21Norm.prox (norm.py:352-364) / L1Norm.prox (norm.py:47-52) with tau' = lam/sigma (ScaleRule).  Substituting,
//   L21:  p - sigma*(p/sigma)*(1 - (lam/sigma)/max(||p||/sigma, lam/sigma))  ==  p * lam / max(||p||, lam)
//   L1 :  p - sigma*soft(p/sigma, lam/sigma)                                 ==  clip(p, -lam, lam)
// i.e. the projections onto the dual balls; evaluated in this closed form (one sqrt + one division per group
// instead of 2G+2 divisions: the unfused arithmetic made the kernels issue-bound).  The results differ from the
// reference's operation order by a few ulps, far inside the 1e-10 / 1e-4 parity tolerance.
// scale of the projection onto the l2 ball of radius lam:  lam / max(||p||, lam) == min(1, lam / ||p||).
// fp32 on the device: one MUFU.RSQ instead of an IEEE sqrt + an IEEE division (each a ~10-instruction sequence with
// a slow-path branch); <= 2 ulp from the exact value, far inside the fp32 tolerance (1e-4).  fp64 stays exact.
PXB_HD float pxb_l21_scale(float nn, float lam) {
#if defined(__CUDA_ARCH__)
    return fminf(1.0f, lam * rsqrtf(nn));
#else
    const float nrm = sqrtf(nn);
    return lam / (nrm > lam ? nrm : lam);
#endif
}
PXB_HD double pxb_l21_scale(double nn, double lam) {
    const double nrm = sqrt(nn);
    return lam / (nrm > lam ? nrm : lam);
}

template <class T>
PXB_HD void pxb_dual_prox_group(int kind, int G, T lam, T sigma, T* p) {
    (void)sigma;
    if (kind == PXB_DUAL_L21) {
        T nn = T(0);
        for (int k = 0; k < G; ++k) nn += p[k] * p[k];
        const T sc = pxb_l21_scale(nn, lam);
        for (int k = 0; k < G; ++k) p[k] = p[k] * sc;
    } else if (kind == PXB_DUAL_L1) {
        for (int k = 0; k < G; ++k) p[k] = p[k] < -lam ? -lam : (p[k] > lam ? lam : p[k]);
    }
}

// gradient of the smooth term at one voxel (pointwise kinds)
template <class T>
PXB_HD T pxb_fgrad_at(const pxb_fterm& f, T x, int64_t lin) {
    if (f.kind == PXB_F_SQL2) {
        const T sh = f.shift ? ((const T*)f.shift)[lin % f.shift_period] : T(0);
        return (x + sh) * T(2 * f.alpha);  // SquaredL2Norm.grad (norm.py:96-98) o ArgShift, ScaleRule
    }
    if (f.kind == PXB_F_GRADARR) return ((const T*)f.garr)[lin];
    return T(0);
}

// ---------------------------------------------------------------------------------------------
// Fused half-iterations at one voxel (see include/pyxu_b200.h for the algebra + citations).
// Returns through references; `lin` = linear index (b, i0, i1, i2) of the voxel.
// ---------------------------------------------------------------------------------------------
template <class T>
PXB_HD void pxb_primal_at(int algo, const pxb_pds_params& P, T ktz, T xu_old, int64_t lin,
                          T& xu_new, T& x_out, T& w_out) {
    const T tau = T(P.tau), rho = T(P.rho);
    if (algo == PXB_PD3O) {
        const T x = pxb_prox_eval<T>(P.g.kind, T(P.g.p0), T(P.g.p1), xu_old - tau * ktz, tau);
        const T ut = x - tau * pxb_fgrad_at<T>(P.f, x, lin);
        w_out = x + ut - xu_old;
        xu_new = (T(1) - rho) * xu_old + rho * ut;
        x_out = x;
    } else {
        const T v = xu_old - tau * pxb_fgrad_at<T>(P.f, xu_old, lin) - tau * ktz;
        const T xt = pxb_prox_eval<T>(P.g.kind, T(P.g.p0), T(P.g.p1), v, tau);
        w_out = T(2) * xt - xu_old;
        xu_new = rho * xt + (T(1) - rho) * xu_old;
        x_out = xu_new;
    }
}

// ---------------------------------------------------------------------------------------------
// Per-voxel kernel bodies.  (b, i0, i1, i2) is the voxel; array base pointers address owned plane 0
// of batch item 0.  `vol` = elements per component incl. halo planes = (n0 + 2*halo)*n1*n2.
// ---------------------------------------------------------------------------------------------
PXB_HD int64_t pxb_vol(const PxbGeom& g, const pxb_slab& s) {
    return (int64_t)(s.plane_alloc > 0 ? s.plane_alloc : g.n0 + 2 * s.halo) * g.s0;
}

template <class T>
PXB_HD void pxb_body_stencil(const pxb_stencil_desc& d, const PxbGeom& g, bool adjoint, const T* __restrict__ in,
                             T* __restrict__ out, int64_t b, int i0, int i1, int i2) {
    const int64_t vol = pxb_vol(g, d.slab);
    const T* img = in + b * vol;
    const int64_t o = b * vol + (int64_t)i0 * g.s0 + (int64_t)i1 * g.s1 + i2;
    out[o] = adjoint ? pxb_stencil_adj_at<T>(d, g, img, (const T*)d.coef, i0, i1, i2)
                     : pxb_stencil_at<T>(d, g, img, (const T*)d.coef, i0, i1, i2);
}

template <class T>
PXB_HD void pxb_body_grad_apply(const pxb_grad_desc& d, const PxbGeom& g, const T* __restrict__ x, T* __restrict__ z,
                                int64_t b, int i0, int i1, int i2) {
    const int64_t vol = pxb_vol(g, d.slab);
    const int64_t v = (int64_t)i0 * g.s0 + (int64_t)i1 * g.s1 + i2;
    for (int k = 0; k < d.ndir; ++k) z[(b * d.ndir + k) * vol + v] = pxb_grad_dir_at<T>(d, g, k, x + b * vol, i0, i1, i2);
}

template <class T>
PXB_HD void pxb_body_grad_adjoint(const pxb_grad_desc& d, const PxbGeom& g, const T* __restrict__ z, T* __restrict__ x,
                                  int64_t b, int i0, int i1, int i2) {
    const int64_t vol = pxb_vol(g, d.slab);
    const int64_t v = (int64_t)i0 * g.s0 + (int64_t)i1 * g.s1 + i2;
    x[b * vol + v] = pxb_grad_adj_at<T>(d, g, z + b * d.ndir * vol, vol, i0, i1, i2);
}

// primal half-step; n0/n1 accumulate sum (x_new - x_old)^2, sum x_old^2 when `want_norms`.
template <class T>
PXB_HD void pxb_body_primal(int algo, const pxb_grad_desc& d, const PxbGeom& g, const pxb_pds_params& P,
                            T* __restrict__ xu, const T* __restrict__ z, const T* __restrict__ ktz, T* __restrict__ x_out,
                            T* __restrict__ w, bool want_norms, double& n0, double& n1, int64_t b, int i0, int i1, int i2) {
    const int64_t vol = pxb_vol(g, d.slab);
    const int64_t lin = b * vol + (int64_t)i0 * g.s0 + (int64_t)i1 * g.s1 + i2;
    T kz = T(0);
    if (ktz) kz = ktz[lin];
    else if (P.hkind != PXB_DUAL_NONE) kz = pxb_grad_adj_at<T>(d, g, z + b * d.ndir * vol, vol, i0, i1, i2);
    const T old = xu[lin];
    T xu_new, xo, wo;
    pxb_primal_at<T>(algo, P, kz, old, lin, xu_new, xo, wo);
    if (want_norms) {
        const T xprev = (algo == PXB_PD3O) ? x_out[lin] : old;
        const double dd = (double)xo - (double)xprev;
        n0 += dd * dd;
        n1 += (double)xprev * (double)xprev;
    }
    xu[lin] = xu_new;
    w[lin] = wo;
    if (algo == PXB_PD3O) x_out[lin] = xo;
}

// dual half-step with in-kernel K w.
template <class T>
PXB_HD void pxb_body_dual(const pxb_grad_desc& d, const PxbGeom& g, const pxb_pds_params& P, const T* __restrict__ w,
                          T* __restrict__ z, bool want_norms, double& n0, double& n1, int64_t b, int i0, int i1, int i2) {
    const int64_t vol = pxb_vol(g, d.slab);
    const int64_t v = (int64_t)i0 * g.s0 + (int64_t)i1 * g.s1 + i2;
    T zo[PXB_MAX_DIRS], p[PXB_MAX_DIRS];
    const T sigma = T(P.sigma), rho = T(P.rho);
    for (int k = 0; k < d.ndir; ++k) {
        zo[k] = z[(b * d.ndir + k) * vol + v];
        p[k] = zo[k] + sigma * pxb_grad_dir_at<T>(d, g, k, w + b * vol, i0, i1, i2);
    }
    pxb_dual_prox_group<T>(P.hkind, d.ndir, T(P.lam), sigma, p);
    for (int k = 0; k < d.ndir; ++k) {
        const T zn = (T(1) - rho) * zo[k] + rho * p[k];
        if (want_norms) {
            const double dd = (double)zn - (double)zo[k];
            n0 += dd * dd;
            n1 += (double)zo[k] * (double)zo[k];
        }
        z[(b * d.ndir + k) * vol + v] = zn;
    }
}

// ---------------------------------------------------------------------------------------------
// Pad over the two trailing axes and its transpose (reference: pad.py:236-375), per sample.  See include/pyxu_b200.h.
// ---------------------------------------------------------------------------------------------
// cell (r, c) of the padded image `img`
template <class T>
PXB_HD T pxb_pad2d_at(const pxb_pad2d_desc& d, const T* __restrict__ in, int64_t img, int r, int c) {
    const int n1 = (int)d.shape[0], n2 = (int)d.shape[1];
    const int e1 = r - d.org[0], e2 = c - d.org[1];
    if (e1 < -d.lo[0] || e1 >= n1 + d.hi[0] || e2 < -d.lo[1] || e2 >= n2 + d.hi[1]) return T(0);  // filler beyond the padded extent
    const int j1 = pxb_bmap(e1, n1, d.mode[0]), j2 = pxb_bmap(e2, n2, d.mode[1]);
    if (j1 == PXB_NOSRC || j2 == PXB_NOSRC) return T(0);
    return in[(img * n1 + j1) * (int64_t)n2 + j2];
}

// sample (t1, t2) of Pad^T ext: every cell of the padded extent that was copied from it
template <class T>
PXB_HD T pxb_pad2d_adj_at(const pxb_pad2d_desc& d, const T* __restrict__ ext, int64_t img, int t1, int t2) {
    const int n1 = (int)d.shape[0], n2 = (int)d.shape[1];
    const int64_t m2 = d.ext_shape[1];
    const T* __restrict__ e = ext + img * d.ext_shape[0] * m2;
    PxbPre P1, P2;
    pxb_preimage(t1, n1, d.mode[0], pxb_imax(d.lo[0], d.hi[0]), 0, 0, P1);
    pxb_preimage(t2, n2, d.mode[1], pxb_imax(d.lo[1], d.hi[1]), 0, 0, P2);
    T acc = T(0);
    for (int a = 0; a < P1.cnt; ++a) {
        const int r_lo = pxb_imax(P1.lo[a], -d.lo[0]), r_hi = pxb_imin(P1.hi[a], n1 + d.hi[0] - 1);
        for (int r = r_lo; r <= r_hi; ++r)
            for (int b = 0; b < P2.cnt; ++b) {
                const int c_lo = pxb_imax(P2.lo[b], -d.lo[1]), c_hi = pxb_imin(P2.hi[b], n2 + d.hi[1] - 1);
                for (int c = c_lo; c <= c_hi; ++c) acc += e[(int64_t)(r + d.org[0]) * m2 + c + d.org[1]];
            }
    }
    return acc;
}

// dual update from a precomputed t = K w, arbitrary group size (outer, group, inner); same closed forms as above.
template <class T>
PXB_HD void pxb_body_dual_update(int kind, int64_t group, int64_t inner, T lam, T sigma, T rho, T* __restrict__ z,
                                 const T* __restrict__ t, bool want_norms, double& n0, double& n1, int64_t o, int64_t i) {
    const int64_t base = o * group * inner + i;
    T sc = T(1);
    if (kind == PXB_DUAL_L21) {
        T nn = T(0);
        for (int64_t k = 0; k < group; ++k) {
            const T p = z[base + k * inner] + sigma * t[base + k * inner];
            nn += p * p;
        }
        const T nrm = sqrt(nn);
        sc = lam / (nrm > lam ? nrm : lam);
    }
    for (int64_t k = 0; k < group; ++k) {
        const T zo = z[base + k * inner];
        const T p = zo + sigma * t[base + k * inner];
        T pr;
        if (kind == PXB_DUAL_L21) pr = p * sc;
        else if (kind == PXB_DUAL_L1) pr = p < -lam ? -lam : (p > lam ? lam : p);
        else pr = p;
        const T zn = (T(1) - rho) * zo + rho * pr;
        if (want_norms) {
            const double dd = (double)zn - (double)zo;
            n0 += dd * dd;
            n1 += (double)zo * (double)zo;
        }
        z[base + k * inner] = zn;
    }
}

// out = prox_{tau*lam*L21}(x) on (outer, group, inner)   (norm.py:352-364)
template <class T>
PXB_HD void pxb_body_prox_l21(int64_t group, int64_t inner, T lam, T tau, const T* __restrict__ x, T* __restrict__ out,
                              int64_t o, int64_t i) {
    const int64_t base = o * group * inner + i;
    const T t = tau * lam;
    T nn = T(0);
    for (int64_t k = 0; k < group; ++k) { const T a = x[base + k * inner]; nn += a * a; }
    const T nrm = sqrt(nn);
    const T sc = T(1) - t / (nrm > t ? nrm : t);
    for (int64_t k = 0; k < group; ++k) out[base + k * inner] = x[base + k * inner] * sc;
}

template <class T>
PXB_HD T pxb_lincomb_at(T a, const T* x, T b, const T* y, int64_t ny, T c, const T* z, int64_t nz, int64_t i) {
    T v = a * x[i];
    if (y) v += b * y[ny ? i % ny : i];
    if (z) v += c * z[nz ? i % nz : i];
    return v;
}
