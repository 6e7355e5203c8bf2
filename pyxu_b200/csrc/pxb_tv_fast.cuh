// pxb_tv_fast.cuh -- specialised bodies of the two fused half-iterations for the dominant case:
//   K = full Gradient stack of the last NDIR axes (direction k acts along axis 3-NDIR+k),
//   every derivative kernel confined to offsets {-1, 0, +1} (first-order forward / backward / central FD).
//
// One thread owns VEC consecutive voxels of a row (16 bytes: 4 fp32 / 2 fp64): all row-aligned operands move
// as 128-bit vector loads/stores; the +-1 neighbours along the row are one extra scalar (L1-resident) load.
// Rows at a domain face whose mode is not 'constant' -- where the reference's boundary extension folds samples
// back -- are delegated, per voxel, to the generic bodies of pxb_core.cuh, so every mode stays exact.
//
// Like pxb_core.cuh these bodies are __host__ __device__: tests/emu runs them on the CPU.
#pragma once
#include "pxb_core.cuh"

struct PxbTvCoef {  // per direction k: taps at offsets -1 / 0 / +1 (0.0 when absent)
    double cm[PXB_MAX_DIRS], c0[PXB_MAX_DIRS], cp[PXB_MAX_DIRS];
};

// host + device: can `d` use the fast bodies?
PXB_HD bool pxb_tv_fast_coefs(const pxb_grad_desc& d, PxbTvCoef& c) {
    if (d.ndir < 1 || d.ndir > 3) return false;
    for (int k = 0; k < d.ndir; ++k) {
        if (d.axis[k] != 3 - d.ndir + k) return false;
        c.cm[k] = c.c0[k] = c.cp[k] = 0.0;
        for (int q = 0; q < d.ntap[k]; ++q) {
            const int o = q - d.center[k];
            if (o == -1) c.cm[k] = d.coef[k][q];
            else if (o == 0) c.c0[k] = d.coef[k][q];
            else if (o == 1) c.cp[k] = d.coef[k][q];
            else return false;
        }
    }
    return true;
}

template <class T, int VEC>
struct PxbVec {
    T v[VEC];
};

template <class T, int VEC>
PXB_HD PxbVec<T, VEC> pxb_vload(const T* __restrict__ p) {
    PxbVec<T, VEC> r;
#if defined(__CUDA_ARCH__)
    if (VEC * sizeof(T) == 16) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        *reinterpret_cast<float4*>(r.v) = t;
    } else if (VEC * sizeof(T) == 8) {
        const float2 t = *reinterpret_cast<const float2*>(p);
        *reinterpret_cast<float2*>(r.v) = t;
    } else
#endif
    {
        for (int j = 0; j < VEC; ++j) r.v[j] = p[j];
    }
    return r;
}

template <class T, int VEC>
PXB_HD void pxb_vstore(T* __restrict__ p, const PxbVec<T, VEC>& r) {
#if defined(__CUDA_ARCH__)
    if (VEC * sizeof(T) == 16) {
        *reinterpret_cast<float4*>(p) = *reinterpret_cast<const float4*>(r.v);
    } else if (VEC * sizeof(T) == 8) {
        *reinterpret_cast<float2*>(p) = *reinterpret_cast<const float2*>(r.v);
    } else
#endif
    {
        for (int j = 0; j < VEC; ++j) p[j] = r.v[j];
    }
}

// Which of this vector's voxels need the generic (folding-aware) body?  A face along axis a matters only when
// its mode is not 'constant' (and, for axis 0, the side is not an open slab cut).  With radius-1 taps the
// boundary extension folds samples onto the face voxel itself (wrap / symmetric / edge) or onto its neighbour
// ('reflect': m(-1) = 1, m(n) = n-2), hence a band of two voxels per face.
template <int NDIR>
PXB_HD bool pxb_tv_needs_generic(const pxb_grad_desc& d, const PxbGeom& g, int i0, int i1, int i2, int vec) {
    if (NDIR >= 3 && d.mode[0] != PXB_CONSTANT)
        if ((i0 <= 1 && !d.slab.open_lo) || (i0 >= g.n0 - 2 && !d.slab.open_hi)) return true;
    if (NDIR >= 2 && d.mode[1] != PXB_CONSTANT)
        if (i1 <= 1 || i1 >= g.n1 - 2) return true;
    if (d.mode[2] != PXB_CONSTANT)
        if (i2 <= 1 || i2 + vec >= g.n2 - 1) return true;
    return false;
}

// shift value for the data term at linear index lin = b*vol + v  (see pxb_fterm)
template <class T>
PXB_HD T pxb_shift_at(const pxb_fterm& f, int64_t lin, int64_t v, int64_t vol) {
    const T* s = (const T*)f.shift;
    if (!s) return T(0);
    if (f.shift_period == 1) return s[0];
    if (f.shift_period == vol) return s[v];
    if (lin < f.shift_period) return s[lin];
    return s[lin % f.shift_period];
}

// ---------------------------------------------------------------------------------------------------------
// primal half-step for VEC voxels (b, i0, i1, i2 .. i2+VEC)
// ---------------------------------------------------------------------------------------------------------
template <class T, int NDIR, int VEC>
PXB_HD void pxb_tv_primal_vec(int algo, const pxb_grad_desc& d, const PxbGeom& g, const PxbTvCoef& cf, const pxb_pds_params& P,
                              T* __restrict__ xu, const T* __restrict__ z, T* __restrict__ x_out, T* __restrict__ w,
                              bool want_norms, double& n0, double& n1, int64_t b, int i0, int i1, int i2) {
    if (pxb_tv_needs_generic<NDIR>(d, g, i0, i1, i2, VEC)) {
        for (int j = 0; j < VEC; ++j)
            pxb_body_primal<T>(algo, d, g, P, xu, z, (const T*)nullptr, x_out, w, want_norms, n0, n1, b, i0, i1, i2 + j);
        return;
    }
    const int64_t vol = pxb_vol(g, d.slab);
    const int64_t v = (int64_t)i0 * g.s0 + (int64_t)i1 * g.s1 + i2;
    const int64_t lin = b * vol + v;
    const T* __restrict__ zb = z + b * NDIR * vol + v;
    T kz[VEC];
    for (int j = 0; j < VEC; ++j) kz[j] = T(0);
    // (K_k^T z)[s] = cm*z_k[s+st] + c0*z_k[s] + cp*z_k[s-st]   (rows outside the domain contribute nothing)
    for (int k = 0; k < NDIR; ++k) {
        const int ax = 3 - NDIR + k;
        const T cm = T(cf.cm[k]), c0 = T(cf.c0[k]), cp = T(cf.cp[k]);
        const T* __restrict__ zk = zb + k * vol;
        if (ax == 2) {
            const PxbVec<T, VEC> c = pxb_vload<T, VEC>(zk);
            T lo = T(0), hi = T(0);
            if (cp != T(0) && i2 > 0) lo = zk[-1];
            if (cm != T(0) && i2 + VEC < g.n2) hi = zk[VEC];
            for (int j = 0; j < VEC; ++j) {
                const T up = (j + 1 < VEC) ? c.v[j + 1 < VEC ? j + 1 : 0] : hi;
                const T dn = (j > 0) ? c.v[j > 0 ? j - 1 : 0] : lo;
                T a = T(0);
                if (cm != T(0)) a += cm * up;
                a += c0 * c.v[j];
                if (cp != T(0)) a += cp * dn;
                kz[j] += a;
            }
        } else {
            const int64_t st = ax == 0 ? g.s0 : g.s1;
            const int i = ax == 0 ? i0 : i1, n = ax == 0 ? g.n0 : g.n1;
            const bool has_lo = i > 0 || (ax == 0 && d.slab.open_lo), has_hi = i < n - 1 || (ax == 0 && d.slab.open_hi);
            const PxbVec<T, VEC> c = pxb_vload<T, VEC>(zk);
            PxbVec<T, VEC> up, dn;
            for (int j = 0; j < VEC; ++j) up.v[j] = dn.v[j] = T(0);
            if (cm != T(0) && has_hi) up = pxb_vload<T, VEC>(zk + st);
            if (cp != T(0) && has_lo) dn = pxb_vload<T, VEC>(zk - st);
            for (int j = 0; j < VEC; ++j) {
                T a = T(0);
                if (cm != T(0)) a += cm * up.v[j];
                a += c0 * c.v[j];
                if (cp != T(0)) a += cp * dn.v[j];
                kz[j] += a;
            }
        }
    }
    const PxbVec<T, VEC> old = pxb_vload<T, VEC>(xu + lin);
    PxbVec<T, VEC> xprev;
    if (want_norms && algo == PXB_PD3O) xprev = pxb_vload<T, VEC>(x_out + lin);
    PxbVec<T, VEC> xn, xo, wo;
    for (int j = 0; j < VEC; ++j) {
        // same algebra as pxb_primal_at, with the shift fetched without a modulo on the common layouts
        const T tau = T(P.tau), rho = T(P.rho);
        const T sh = (P.f.kind == PXB_F_SQL2) ? pxb_shift_at<T>(P.f, lin + j, v + j, vol) : T(0);
        if (algo == PXB_PD3O) {
            const T x = pxb_prox_eval<T>(P.g.kind, T(P.g.p0), T(P.g.p1), old.v[j] - tau * kz[j], tau);
            const T gf = (P.f.kind == PXB_F_SQL2) ? (x + sh) * T(2 * P.f.alpha) : T(0);
            const T ut = x - tau * gf;
            wo.v[j] = x + ut - old.v[j];
            xn.v[j] = (T(1) - rho) * old.v[j] + rho * ut;
            xo.v[j] = x;
        } else {
            T gf = T(0);
            if (P.f.kind == PXB_F_SQL2) gf = (old.v[j] + sh) * T(2 * P.f.alpha);
            else if (P.f.kind == PXB_F_GRADARR) gf = ((const T*)P.f.garr)[lin + j];
            const T vv = old.v[j] - tau * gf - tau * kz[j];
            const T xt = pxb_prox_eval<T>(P.g.kind, T(P.g.p0), T(P.g.p1), vv, tau);
            wo.v[j] = T(2) * xt - old.v[j];
            xn.v[j] = rho * xt + (T(1) - rho) * old.v[j];
            xo.v[j] = xn.v[j];
        }
        if (want_norms) {
            const T xp = (algo == PXB_PD3O) ? xprev.v[j] : old.v[j];
            const double dd = (double)xo.v[j] - (double)xp;
            n0 += dd * dd;
            n1 += (double)xp * (double)xp;
        }
    }
    pxb_vstore<T, VEC>(xu + lin, xn);
    pxb_vstore<T, VEC>(w + lin, wo);
    if (algo == PXB_PD3O) pxb_vstore<T, VEC>(x_out + lin, xo);
}

// ---------------------------------------------------------------------------------------------------------
// dual half-step for VEC voxels:  z <- (1-rho) z + rho prox_{sigma h*}(z + sigma K w)
// ---------------------------------------------------------------------------------------------------------
template <class T, int NDIR, int VEC>
PXB_HD void pxb_tv_dual_vec(const pxb_grad_desc& d, const PxbGeom& g, const PxbTvCoef& cf, const pxb_pds_params& P,
                            const T* __restrict__ w, T* __restrict__ z, bool want_norms, double& n0, double& n1, int64_t b,
                            int i0, int i1, int i2) {
    if (pxb_tv_needs_generic<NDIR>(d, g, i0, i1, i2, VEC)) {
        for (int j = 0; j < VEC; ++j) pxb_body_dual<T>(d, g, P, w, z, want_norms, n0, n1, b, i0, i1, i2 + j);
        return;
    }
    const int64_t vol = pxb_vol(g, d.slab);
    const int64_t v = (int64_t)i0 * g.s0 + (int64_t)i1 * g.s1 + i2;
    const T* __restrict__ wb = w + b * vol + v;
    T* __restrict__ zb = z + b * NDIR * vol + v;
    const PxbVec<T, VEC> wc = pxb_vload<T, VEC>(wb);
    T p[NDIR][VEC], zo[NDIR][VEC];
    const T sigma = T(P.sigma), rho = T(P.rho);
    // (K_k w)[s] = cm*w[s-st] + c0*w[s] + cp*w[s+st]
    for (int k = 0; k < NDIR; ++k) {
        const int ax = 3 - NDIR + k;
        const T cm = T(cf.cm[k]), c0 = T(cf.c0[k]), cp = T(cf.cp[k]);
        const PxbVec<T, VEC> zc = pxb_vload<T, VEC>(zb + k * vol);
        T kw[VEC];
        if (ax == 2) {
            T lo = T(0), hi = T(0);
            if (cm != T(0) && i2 > 0) lo = wb[-1];
            if (cp != T(0) && i2 + VEC < g.n2) hi = wb[VEC];
            for (int j = 0; j < VEC; ++j) {
                const T up = (j + 1 < VEC) ? wc.v[j + 1 < VEC ? j + 1 : 0] : hi;
                const T dn = (j > 0) ? wc.v[j > 0 ? j - 1 : 0] : lo;
                T a = T(0);
                if (cm != T(0)) a += cm * dn;
                a += c0 * wc.v[j];
                if (cp != T(0)) a += cp * up;
                kw[j] = a;
            }
        } else {
            const int64_t st = ax == 0 ? g.s0 : g.s1;
            const int i = ax == 0 ? i0 : i1, n = ax == 0 ? g.n0 : g.n1;
            const bool has_lo = i > 0 || (ax == 0 && d.slab.open_lo), has_hi = i < n - 1 || (ax == 0 && d.slab.open_hi);
            PxbVec<T, VEC> up, dn;
            for (int j = 0; j < VEC; ++j) up.v[j] = dn.v[j] = T(0);
            if (cp != T(0) && has_hi) up = pxb_vload<T, VEC>(wb + st);
            if (cm != T(0) && has_lo) dn = pxb_vload<T, VEC>(wb - st);
            for (int j = 0; j < VEC; ++j) {
                T a = T(0);
                if (cm != T(0)) a += cm * dn.v[j];
                a += c0 * wc.v[j];
                if (cp != T(0)) a += cp * up.v[j];
                kw[j] = a;
            }
        }
        for (int j = 0; j < VEC; ++j) {
            zo[k][j] = zc.v[j];
            p[k][j] = zc.v[j] + sigma * kw[j];
        }
    }
    for (int j = 0; j < VEC; ++j) {
        T grp[PXB_MAX_DIRS];
        for (int k = 0; k < NDIR; ++k) grp[k] = p[k][j];
        pxb_dual_prox_group<T>(P.hkind, NDIR, T(P.lam), sigma, grp);
        for (int k = 0; k < NDIR; ++k) {
            const T zn = (T(1) - rho) * zo[k][j] + rho * grp[k];
            if (want_norms) {
                const double dd = (double)zn - (double)zo[k][j];
                n0 += dd * dd;
                n1 += (double)zo[k][j] * (double)zo[k][j];
            }
            p[k][j] = zn;
        }
    }
    for (int k = 0; k < NDIR; ++k) {
        PxbVec<T, VEC> o;
        for (int j = 0; j < VEC; ++j) o.v[j] = p[k][j];
        pxb_vstore<T, VEC>(zb + k * vol, o);
    }
}
