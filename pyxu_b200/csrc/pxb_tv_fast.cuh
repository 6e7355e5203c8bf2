// pxb_tv_fast.cuh -- specialised bodies of the two fused half-iterations for the dominant case:
//   K = full Gradient stack of the last NDIR axes (direction k acts along axis 3-NDIR+k),
//   every derivative kernel confined to offsets {-1, 0, +1} (first-order forward / backward / central FD).
//
// One thread owns VEC consecutive voxels of a row (16 bytes: 4 fp32 / 2 fp64): all row-aligned operands move
// as 128-bit vector loads/stores; the +-1 neighbours along the row are one extra scalar (L1-resident) load.
// Voxels next to a domain face whose mode is not 'constant' -- where the reference's boundary extension folds
// samples back -- are delegated, per voxel, to the generic bodies of pxb_core.cuh, so every mode stays exact.
//
// The first version of these bodies was issue-bound, not HBM-bound (ncu, profiles/r01_a_*: 74 % issue-active,
// ~150 instructions per voxel): all scalars are therefore converted to T and folded on the host (PxbTvP<T>), the
// dual prox is evaluated in its closed form (one sqrt + one division per voxel), and the shift / gradient
// arrays of the data term are vector-loaded.
//
// Like pxb_core.cuh these bodies are __host__ __device__: tests/emu runs them on the CPU.
#pragma once
#include "pxb_core.cuh"

#if defined(__CUDACC__)
#define PXB_NOINLINE __host__ __device__ __noinline__
#else
#define PXB_NOINLINE inline
#endif

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

// how the shift of the data term f = alpha*||x + shift||^2 is addressed
enum { PXB_SHIFT_NONE = 0, PXB_SHIFT_SCALAR = 1, PXB_SHIFT_LIN = 2 /* shift[lin] */, PXB_SHIFT_VOL = 3 /* shift[v] */, PXB_SHIFT_MOD = 4 };

// Everything the fast bodies need, typed and folded on the host.
template <class T>
struct PxbTvP {
    T cm[PXB_MAX_DIRS], c0[PXB_MAX_DIRS], cp[PXB_MAX_DIRS];
    T tau, sigma, rho, one_m_rho, lam, two_alpha, gp0, gp1;
    int rho1;  // rho == 1 (the reference's default): the relaxation (1-rho) old + rho new IS new -- the tiled forms skip its arithmetic
    int gkind, fkind, hkind, shift_mode;
    const T* shift;
    const T* garr;
    int64_t shift_period;
    int n0, n1, n2;
    int64_t s0, s1, vol;  // vol: elements between components / batch items
    int mode[3];
    int open_lo, open_hi;
    // Folding boundary modes, per AXIS (see pxb_tv_fold_kz): the sample whose K^T z gains cp*z[n-1] (fold_hi) / cm*z[0]
    // (fold_lo) on top of the 'constant' arithmetic; PXB_NOSRC when the axis does not fold or the tap is absent.
    int fold_hi[3], fold_lo[3];
    // Packed fp32 forms (pxb_iter_phaseC_f32x2, pxb_tma_w): the folded coefficients as (c, c) pairs, read by FFMA2 / FMUL2 as
    // uniform-register operands straight from the constant bank (computed per call and per thread they were 7 % of the instructions):
    // [0..2] sigma*c0[k], [3..5] sigma*cp[k], [6..8] -tau*c0[k], [9..11] -tau*cp[k], [12] 1 - 2 alpha tau, [13] -2 alpha tau, [14] -1
    alignas(8) T pk[15][2];
    T one;  // 1, opaque to the compiler: `v * one` is an exact copy of v that cannot be issued before v's load has landed (pxb_landed)
};


// ---------------------------------------------------------------------------------------------------------
// Compile-time specialisation of the single-kernel iteration.  -1 / 0 = decided at run time (the generic instance
// handles every case); the other values let the compiler drop the tap tests, the prox switch and the f / h
// dispatch for the configurations the named workloads use (forward differences, L21, shifted squared-l2 data term,
// positivity or no constraint).  ncu on the generic instance: 11 % of the issued instructions were constant-bank
// reloads and 15 % branches / reconvergence barriers, with the kernel issue-bound at 64 % issue utilisation.
// ---------------------------------------------------------------------------------------------------------
enum { PXB_SCHEME_ANY = 0, PXB_SCHEME_FWD = 1 /* every direction: taps at {0, +1}: cm == 0, cp != 0 */ };
template <int SCHEME_, int GK_, int HK_, int FK_>
struct PxbSpec {
    static constexpr int SCHEME = SCHEME_;  // PXB_SCHEME_*
    static constexpr int GK = GK_;          // -1 run time, else pxb_prox_kind
    static constexpr int HK = HK_;          // -1 run time, else pxb_dual_kind
    static constexpr int FK = FK_;          // -1 run time, 1: f = alpha ||x + shift||^2 with a per-voxel shift array,
                                            //  2: grad f handed over as a per-voxel array (CondatVu with a non-local f)
};
using PxbSpecAny = PxbSpec<PXB_SCHEME_ANY, -1, -1, -1>;

template <class S, class T> PXB_HD bool pxb_has_cm(const PxbTvP<T>& q, int k) { return S::SCHEME == PXB_SCHEME_FWD ? false : q.cm[k] != T(0); }
template <class S, class T> PXB_HD bool pxb_has_cp(const PxbTvP<T>& q, int k) { return S::SCHEME == PXB_SCHEME_FWD ? true : q.cp[k] != T(0); }
template <class S, class T> PXB_HD int pxb_gkind(const PxbTvP<T>& q) { return S::GK >= 0 ? S::GK : q.gkind; }
template <class S, class T> PXB_HD int pxb_hkind(const PxbTvP<T>& q) { return S::HK >= 0 ? S::HK : q.hkind; }

template <class T>
PXB_HD void pxb_tv_prepare(const pxb_grad_desc& d, const PxbTvCoef& cf, const pxb_pds_params& P, PxbTvP<T>& q) {
    for (int k = 0; k < PXB_MAX_DIRS; ++k) {
        q.cm[k] = T(cf.cm[k]); q.c0[k] = T(cf.c0[k]); q.cp[k] = T(cf.cp[k]);
    }
    q.tau = T(P.tau); q.sigma = T(P.sigma); q.rho = T(P.rho); q.one_m_rho = T(1) - q.rho; q.rho1 = P.rho == 1.0 ? 1 : 0;
    q.lam = T(P.lam); q.two_alpha = T(2 * P.f.alpha); q.gp0 = T(P.g.p0); q.gp1 = T(P.g.p1);
    {
        T v[15];
        for (int k = 0; k < 3; ++k) {
            v[k] = q.sigma * q.c0[k]; v[3 + k] = q.sigma * q.cp[k];
            v[6 + k] = -q.tau * q.c0[k]; v[9 + k] = -q.tau * q.cp[k];
        }
        v[13] = -q.tau * q.two_alpha; v[12] = T(1) + v[13]; v[14] = T(-1);
        for (int i = 0; i < 15; ++i) q.pk[i][0] = q.pk[i][1] = v[i];
        q.one = T(1);
    }
    q.gkind = P.g.kind; q.fkind = P.f.kind; q.hkind = P.hkind;
    const PxbGeom g = pxb_geom(d.shape);
    q.n0 = g.n0; q.n1 = g.n1; q.n2 = g.n2; q.s0 = g.s0; q.s1 = g.s1;
    q.vol = pxb_vol(g, d.slab);
    for (int a = 0; a < 3; ++a) { q.mode[a] = d.mode[a]; q.fold_hi[a] = q.fold_lo[a] = PXB_NOSRC; }
    q.open_lo = d.slab.open_lo; q.open_hi = d.slab.open_hi;
    for (int k = 0; k < d.ndir && k < PXB_MAX_DIRS; ++k) {  // an open slab side is not a boundary: nothing folds there
        const int ax = 3 - d.ndir + k, n = ax == 0 ? g.n0 : (ax == 1 ? g.n1 : g.n2);
        if (ax < 0 || d.mode[ax] == PXB_CONSTANT) continue;
        if (cf.cp[k] != 0.0 && !(ax == 0 && d.slab.open_hi)) q.fold_hi[ax] = pxb_bmap(n, n, d.mode[ax]);
        if (cf.cm[k] != 0.0 && !(ax == 0 && d.slab.open_lo)) q.fold_lo[ax] = pxb_bmap(-1, n, d.mode[ax]);
    }
    q.shift = (const T*)P.f.shift; q.garr = (const T*)P.f.garr; q.shift_period = P.f.shift_period;
    q.shift_mode = PXB_SHIFT_NONE;
    if (P.f.kind == PXB_F_SQL2 && P.f.shift) {
        const int64_t span = (d.batch - 1) * q.vol + (int64_t)g.n0 * g.s0;  // largest lin + 1
        if (P.f.shift_period == 1) q.shift_mode = PXB_SHIFT_SCALAR;
        else if (P.f.shift_period >= span) q.shift_mode = PXB_SHIFT_LIN;
        else if (P.f.shift_period == q.vol) q.shift_mode = PXB_SHIFT_VOL;
        else q.shift_mode = PXB_SHIFT_MOD;
    }
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

// taps of one direction along the row itself:  out[j] = c_hi * f[j+1] + c_0 * f[j] + c_lo * f[j-1]
// (the vector's own elements serve as neighbours, the two ends come from one scalar load each)
// off_lo / off_hi: where the neighbours of the vector's first / last element live relative to f (-1 / VEC inside the
// row; the sample the boundary map folds them onto at a folding face, see pxb_tv_nbr)
template <class T, int VEC>
PXB_HD void pxb_tv_taps_row(const T* __restrict__ f, const PxbVec<T, VEC>& c, T c_hi, T c_0, T c_lo, bool has_lo, bool has_hi, T* out,
                            int off_lo = -1, int off_hi = VEC) {
    T lo = T(0), hi = T(0);
    if (c_lo != T(0) && has_lo) lo = f[off_lo];
    if (c_hi != T(0) && has_hi) hi = f[off_hi];
    for (int j = 0; j < VEC; ++j) {
        const T up = (j + 1 < VEC) ? c.v[j + 1 < VEC ? j + 1 : 0] : hi;
        const T dn = (j > 0) ? c.v[j > 0 ? j - 1 : 0] : lo;
        T a = c_0 * c.v[j];
        if (c_hi != T(0)) a += c_hi * up;
        if (c_lo != T(0)) a += c_lo * dn;
        out[j] = a;
    }
}

// taps of one direction across rows / planes:  out[j] = c_hi * f[s + d_hi st] + c_0 * f[s] + c_lo * f[s + d_lo st]
// (d_lo, d_hi) = (-1, +1) inside the domain; at a folding face the step to the sample the neighbour folds onto
template <class T, int VEC>
PXB_HD void pxb_tv_taps_col(const T* __restrict__ f, int64_t st, const PxbVec<T, VEC>& c, T c_hi, T c_0, T c_lo, bool has_lo, bool has_hi,
                            T* out, int d_lo = -1, int d_hi = 1) {
    for (int j = 0; j < VEC; ++j) out[j] = c_0 * c.v[j];
    if (c_hi != T(0) && has_hi) {
        const PxbVec<T, VEC> up = pxb_vload<T, VEC>(f + d_hi * st);
        for (int j = 0; j < VEC; ++j) out[j] += c_hi * up.v[j];
    }
    if (c_lo != T(0) && has_lo) {
        const PxbVec<T, VEC> dn = pxb_vload<T, VEC>(f + d_lo * st);
        for (int j = 0; j < VEC; ++j) out[j] += c_lo * dn.v[j];
    }
}

// Neighbours of samples [i, i + w) of a line of length n under K (apply direction): (K w)[s] = cm w[m(s-1)] + c0 w[s] +
// cp w[m(s+1)] with m the boundary map (pad.py:252-302).  d_lo / d_hi: step from sample i to the lower neighbour / from
// sample i + w - 1 to the upper one (-1 / +1 inside the line, the fold at a folding face); has_*: false where the
// neighbour is a zero of the 'constant' extension.  Open slab sides (axis 0) read their ghost planes.
template <class T>
PXB_HD void pxb_tv_nbr(const PxbTvP<T>& q, int ax, int i, int w, int n, int& d_lo, int& d_hi, bool& has_lo, bool& has_hi) {
    d_lo = -1;
    d_hi = 1;
    has_lo = i > 0 || (ax == 0 && q.open_lo);
    has_hi = i + w < n || (ax == 0 && q.open_hi);
    if (q.mode[ax] != PXB_CONSTANT) {
        if (!has_lo) { d_lo = pxb_bmap(-1, n, q.mode[ax]) - i; has_lo = true; }
        if (!has_hi) { d_hi = pxb_bmap(n, n, q.mode[ax]) - (i + w - 1); has_hi = true; }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Folding boundary modes (numpy.pad wrap / reflect / symmetric / edge, pad.py:252-302) for radius-1 taps.
// K_k = Trim o S_k o Pad_k pads axis ax by one sample on the side(s) where S_k has a tap, so
//     (K_k w)[s]   = cm w[m(s-1)] + c0 w[s] + cp w[m(s+1)]             m = the boundary map (pxb_bmap)
//     (K_k^T z)[t] = 'constant' arithmetic + [t == m(n)] cp z[n-1] + [t == m(-1)] cm z[0]
// (the transpose of Pad adds the padded cell n, resp. -1, onto the sample it was copied from, pad.py:307-375; of the
// taps that reach that cell only z[n-1] through cp, resp. z[0] through cm, lie inside the domain).  One sample per
// line and side gains one term: m(n) = 0 | n-2 | n-1 | n-1 and m(-1) = n-1 | 1 | 0 | 0 for wrap | reflect | symmetric |
// edge -- folded on the host into PxbTvP::fold_hi / fold_lo.  The term is added from global memory (one coalesced
// vector load for a row / plane face, one scalar load by one lane for a column face).
//   zimg: component 0 of the batch item, sample (0, 0, 0);  the W samples (i0, i1, i2..) lie inside the domain.
// ---------------------------------------------------------------------------------------------------------
// (The tests come first and everything else -- addresses included -- sits inside the taken branches: ncu on the 512^3 reflect
//  instance showed the unconditional 64-bit address arithmetic of the first version as 125 M of the 182 M instructions the MODES
//  instance executed on top of the 'constant' one.)  LO = false drops the fold_lo terms at compile time (forward differences).
template <class T, int W, int NDIR, bool LO = true>
PXB_HD void pxb_tv_fold_kz(const PxbTvP<T>& q, const T* __restrict__ zimg, int i0, int i1, int i2, T* kz) {
    for (int k = 0; k < NDIR; ++k) {
        const int ax = 3 - NDIR + k;
        if (ax == 2) {
            const int th = q.fold_hi[2], tl = LO ? q.fold_lo[2] : PXB_NOSRC;
            const bool hi = th >= i2 && th < i2 + W, lo = LO && tl >= i2 && tl < i2 + W;
            if (hi || lo) {
                const T* __restrict__ row = zimg + k * q.vol + (int64_t)i0 * q.s0 + (int64_t)i1 * q.s1;
                if (hi) {
                    const T zf = row[q.n2 - 1];
                    for (int j = 0; j < W; ++j)
                        if (i2 + j == th) kz[j] += q.cp[k] * zf;
                }
                if (lo) {
                    const T zf = row[0];
                    for (int j = 0; j < W; ++j)
                        if (i2 + j == tl) kz[j] += q.cm[k] * zf;
                }
            }
        } else {
            const int i = ax == 0 ? i0 : i1;
            const bool hi = i == q.fold_hi[ax], lo = LO && i == q.fold_lo[ax];
            if (hi || lo) {
                const int n = ax == 0 ? q.n0 : q.n1;
                const int64_t st = ax == 0 ? q.s0 : q.s1;
                const T* __restrict__ line = zimg + k * q.vol + (int64_t)i0 * q.s0 + (int64_t)i1 * q.s1 + i2 - (int64_t)i * st;
                if (hi) {
                    const PxbVec<T, W> f = pxb_vload<T, W>(line + (int64_t)(n - 1) * st);
                    for (int j = 0; j < W; ++j) kz[j] += q.cp[k] * f.v[j];
                }
                if (lo) {
                    const PxbVec<T, W> f = pxb_vload<T, W>(line);
                    for (int j = 0; j < W; ++j) kz[j] += q.cm[k] * f.v[j];
                }
            }
        }
    }
}

// w = the point K is applied to in the dual half-step, for W in-domain samples starting at (b, i0, i1, i2), straight
// from global memory with every boundary mode honoured.  The tiled single-kernel forms call it for the cells of
// their w tiles that lie one step outside the domain (K w reads there w at the sample the boundary map folds the cell
// onto); out of line: a vanishing fraction of the cells takes it, its registers must not weigh on the tiled path.
template <class T, int W, int NDIR, int ALGO>
PXB_NOINLINE void pxb_tv_w_global(const PxbTvP<T>& q, const T* __restrict__ xu, const T* __restrict__ z, int64_t b, int i0, int i1, int i2, T* wv) {
#if defined(PXB_EMU_COUNT_W_GLOBAL)  // tests/emu only: how many cells took this path
    PXB_EMU_COUNT_W_GLOBAL += W;
#endif
    const int64_t v = (int64_t)i0 * q.s0 + (int64_t)i1 * q.s1 + i2;
    const int64_t lin = b * q.vol + v;
    const T* __restrict__ zimg = z + b * NDIR * q.vol;
    T kz[W];
    for (int j = 0; j < W; ++j) kz[j] = T(0);
    for (int k = 0; k < NDIR; ++k) {
        const int ax = 3 - NDIR + k;
        const T* __restrict__ zk = zimg + k * q.vol + v;
        const PxbVec<T, W> c = pxb_vload<T, W>(zk);
        T t[W];
        if (ax == 2) {
            pxb_tv_taps_row<T, W>(zk, c, q.cm[k], q.c0[k], q.cp[k], i2 > 0, i2 + W < q.n2, t);
        } else {
            const int i = ax == 0 ? i0 : i1, n = ax == 0 ? q.n0 : q.n1;
            const bool has_lo = i > 0 || (ax == 0 && q.open_lo), has_hi = i < n - 1 || (ax == 0 && q.open_hi);  // slab cuts read their ghost planes
            pxb_tv_taps_col<T, W>(zk, ax == 0 ? q.s0 : q.s1, c, q.cm[k], q.c0[k], q.cp[k], has_lo, has_hi, t);
        }
        for (int j = 0; j < W; ++j) kz[j] += t[j];
    }
    pxb_tv_fold_kz<T, W, NDIR>(q, zimg, i0, i1, i2, kz);
    const PxbVec<T, W> old = pxb_vload<T, W>(xu + lin);
    PxbVec<T, W> sh;
    for (int j = 0; j < W; ++j) sh.v[j] = T(0);
    if (q.fkind == PXB_F_SQL2) {
        if (q.shift_mode == PXB_SHIFT_LIN) sh = pxb_vload<T, W>(q.shift + lin);
        else if (q.shift_mode == PXB_SHIFT_VOL) sh = pxb_vload<T, W>(q.shift + v);
        else if (q.shift_mode == PXB_SHIFT_SCALAR) { for (int j = 0; j < W; ++j) sh.v[j] = q.shift[0]; }
        else if (q.shift_mode == PXB_SHIFT_MOD) { for (int j = 0; j < W; ++j) sh.v[j] = q.shift[(lin + j) % q.shift_period]; }
    } else if (q.fkind == PXB_F_GRADARR) {
        sh = pxb_vload<T, W>(q.garr + lin);
    }
    for (int j = 0; j < W; ++j) {
        if (ALGO == PXB_PD3O) {
            const T x = pxb_prox_eval<T>(q.gkind, q.gp0, q.gp1, old.v[j] - q.tau * kz[j], q.tau);
            const T gf = (q.fkind == PXB_F_SQL2) ? (x + sh.v[j]) * q.two_alpha : T(0);
            wv[j] = x + (x - q.tau * gf) - old.v[j];
        } else {
            T gf = T(0);
            if (q.fkind == PXB_F_SQL2) gf = (old.v[j] + sh.v[j]) * q.two_alpha;
            else if (q.fkind == PXB_F_GRADARR) gf = sh.v[j];
            const T xt = pxb_prox_eval<T>(q.gkind, q.gp0, q.gp1, old.v[j] - q.tau * gf - q.tau * kz[j], q.tau);
            wv[j] = T(2) * xt - old.v[j];
        }
    }
}

// The W cells starting at (i0, i1, i2) of a w tile, not all inside the domain: what K w reads there.  A cell one step
// outside along exactly one folding axis holds w at the sample the boundary map sends it to; nothing reads the others
// ('constant' axes read zeros; corners and cells further out are never touched by an axis-aligned radius-1 stencil).
// Row vectors are W-aligned and n2 is a multiple of W, so along the columns only cell 0 of the vector can be "n2".
template <class T, int W, int NDIR, int ALGO>
PXB_HD void pxb_tv_w_outside(const PxbTvP<T>& q, const T* __restrict__ xu, const T* __restrict__ z, int64_t b, int i0, int i1, int i2, T* wv) {
    for (int j = 0; j < W; ++j) wv[j] = T(0);
    // (the ghost planes of an open slab side are not outside: callers only come here for them at in-plane cells nobody reads)
    const bool o0 = NDIR == 3 && ((i0 < 0 && !q.open_lo) || (i0 >= q.n0 && !q.open_hi)), o1 = i1 < 0 || i1 >= q.n1, o2 = i2 < 0 || i2 >= q.n2;
    if (NDIR == 3 && !o0 && (i0 < 0 || i0 >= q.n0)) return;
    if ((o0 ? 1 : 0) + (o1 ? 1 : 0) + (o2 ? 1 : 0) != 1) return;
    if (NDIR == 2 && (i0 < 0 || i0 >= q.n0)) return;
    if (o2) {
        if ((i2 != -1 && i2 != q.n2) || q.mode[2] == PXB_CONSTANT) return;
        T w1[1];
        pxb_tv_w_global<T, 1, NDIR, ALGO>(q, xu, z, b, i0, i1, pxb_bmap(i2, q.n2, q.mode[2]), w1);
        wv[0] = w1[0];
    } else if (o1) {
        if ((i1 != -1 && i1 != q.n1) || q.mode[1] == PXB_CONSTANT) return;
        pxb_tv_w_global<T, W, NDIR, ALGO>(q, xu, z, b, i0, pxb_bmap(i1, q.n1, q.mode[1]), i2, wv);
    } else {
        if ((i0 != -1 && i0 != q.n0) || q.mode[0] == PXB_CONSTANT) return;
        pxb_tv_w_global<T, W, NDIR, ALGO>(q, xu, z, b, pxb_bmap(i0, q.n0, q.mode[0]), i1, i2, wv);
    }
}

// ---------------------------------------------------------------------------------------------------------
// primal half-step for VEC voxels (b, i0, i1, i2 .. i2+VEC).  nrm[0..1] += RelError[x] partial sums.
// ---------------------------------------------------------------------------------------------------------
template <class T, int NDIR, int VEC, int ALGO, bool NORMS>
PXB_HD void pxb_tv_primal_vec(const PxbTvP<T>& q, const pxb_grad_desc& d, const pxb_pds_params& P, T* __restrict__ xu,
                              const T* __restrict__ z, T* __restrict__ x_out, T* __restrict__ w, double* nrm, int64_t b, int i0,
                              int i1, int i2) {
    const int64_t v = (int64_t)i0 * q.s0 + (int64_t)i1 * q.s1 + i2;
    const int64_t lin = b * q.vol + v;
    const T* __restrict__ zb = z + b * NDIR * q.vol + v;
    T kz[VEC];
    for (int j = 0; j < VEC; ++j) kz[j] = T(0);
    // (K_k^T z)[s] = cm*z_k[s+st] + c0*z_k[s] + cp*z_k[s-st]   (rows outside the domain contribute nothing; folding
    // faces add their term below)
    for (int k = 0; k < NDIR; ++k) {
        const int ax = 3 - NDIR + k;
        const T* __restrict__ zk = zb + k * q.vol;
        const PxbVec<T, VEC> c = pxb_vload<T, VEC>(zk);
        T t[VEC];
        if (ax == 2) {
            pxb_tv_taps_row<T, VEC>(zk, c, q.cm[k], q.c0[k], q.cp[k], i2 > 0, i2 + VEC < q.n2, t);
        } else {
            const int i = ax == 0 ? i0 : i1, n = ax == 0 ? q.n0 : q.n1;
            const bool has_lo = i > 0 || (ax == 0 && q.open_lo), has_hi = i < n - 1 || (ax == 0 && q.open_hi);
            pxb_tv_taps_col<T, VEC>(zk, ax == 0 ? q.s0 : q.s1, c, q.cm[k], q.c0[k], q.cp[k], has_lo, has_hi, t);
        }
        for (int j = 0; j < VEC; ++j) kz[j] += t[j];
    }
    pxb_tv_fold_kz<T, VEC, NDIR>(q, z + b * NDIR * q.vol, i0, i1, i2, kz);
    const PxbVec<T, VEC> old = pxb_vload<T, VEC>(xu + lin);
    PxbVec<T, VEC> sh;
    for (int j = 0; j < VEC; ++j) sh.v[j] = T(0);
    if (q.fkind == PXB_F_SQL2) {
        if (q.shift_mode == PXB_SHIFT_LIN) sh = pxb_vload<T, VEC>(q.shift + lin);
        else if (q.shift_mode == PXB_SHIFT_VOL) sh = pxb_vload<T, VEC>(q.shift + v);
        else if (q.shift_mode == PXB_SHIFT_SCALAR) { for (int j = 0; j < VEC; ++j) sh.v[j] = q.shift[0]; }
        else if (q.shift_mode == PXB_SHIFT_MOD) { for (int j = 0; j < VEC; ++j) sh.v[j] = q.shift[(lin + j) % q.shift_period]; }
    }
    PxbVec<T, VEC> xn, xo, wo;
    T a0 = T(0), a1 = T(0);  // one vector's RelError partial sums in the working precision, widened once
    if (ALGO == PXB_PD3O) {
        PxbVec<T, VEC> xprev;
        if (NORMS) xprev = pxb_vload<T, VEC>(x_out + lin);
        for (int j = 0; j < VEC; ++j) {
            const T x = pxb_prox_eval<T>(q.gkind, q.gp0, q.gp1, old.v[j] - q.tau * kz[j], q.tau);
            const T gf = (q.fkind == PXB_F_SQL2) ? (x + sh.v[j]) * q.two_alpha : T(0);
            const T ut = x - q.tau * gf;
            wo.v[j] = x + ut - old.v[j];
            xn.v[j] = q.one_m_rho * old.v[j] + q.rho * ut;
            xo.v[j] = x;
            if (NORMS) {
                const T dd = x - xprev.v[j];
                a0 += dd * dd;
                a1 += xprev.v[j] * xprev.v[j];
            }
        }
        pxb_vstore<T, VEC>(x_out + lin, xo);
    } else {
        PxbVec<T, VEC> ga;
        if (q.fkind == PXB_F_GRADARR) ga = pxb_vload<T, VEC>(q.garr + lin);
        for (int j = 0; j < VEC; ++j) {
            T gf = T(0);
            if (q.fkind == PXB_F_SQL2) gf = (old.v[j] + sh.v[j]) * q.two_alpha;
            else if (q.fkind == PXB_F_GRADARR) gf = ga.v[j];
            const T vv = old.v[j] - q.tau * gf - q.tau * kz[j];
            const T xt = pxb_prox_eval<T>(q.gkind, q.gp0, q.gp1, vv, q.tau);
            wo.v[j] = T(2) * xt - old.v[j];
            xn.v[j] = q.rho * xt + q.one_m_rho * old.v[j];
            if (NORMS) {
                const T dd = xn.v[j] - old.v[j];
                a0 += dd * dd;
                a1 += old.v[j] * old.v[j];
            }
        }
    }
    pxb_vstore<T, VEC>(xu + lin, xn);
    pxb_vstore<T, VEC>(w + lin, wo);
    if (NORMS) { nrm[0] += (double)a0; nrm[1] += (double)a1; }
}

// ---------------------------------------------------------------------------------------------------------
// dual half-step for VEC voxels:  z <- (1-rho) z + rho prox_{sigma h*}(z + sigma K w)
// ---------------------------------------------------------------------------------------------------------
template <class T, int NDIR, int VEC, bool NORMS>
PXB_HD void pxb_tv_dual_vec(const PxbTvP<T>& q, const pxb_grad_desc& d, const pxb_pds_params& P, const T* __restrict__ w,
                            T* __restrict__ z, double* nrm, int64_t b, int i0, int i1, int i2) {
    const int64_t v = (int64_t)i0 * q.s0 + (int64_t)i1 * q.s1 + i2;
    const T* __restrict__ wb = w + b * q.vol + v;
    T* __restrict__ zb = z + b * NDIR * q.vol + v;
    const PxbVec<T, VEC> wc = pxb_vload<T, VEC>(wb);
    T p[NDIR][VEC], zo[NDIR][VEC];
    // (K_k w)[s] = cm*w[s-st] + c0*w[s] + cp*w[s+st]
    for (int k = 0; k < NDIR; ++k) {
        const int ax = 3 - NDIR + k;
        const PxbVec<T, VEC> zc = pxb_vload<T, VEC>(zb + k * q.vol);
        T kw[VEC];
        int d_lo, d_hi;
        bool has_lo, has_hi;
        if (ax == 2) {
            pxb_tv_nbr<T>(q, 2, i2, VEC, q.n2, d_lo, d_hi, has_lo, has_hi);
            pxb_tv_taps_row<T, VEC>(wb, wc, q.cp[k], q.c0[k], q.cm[k], has_lo, has_hi, kw, d_lo, VEC - 1 + d_hi);
        } else {
            pxb_tv_nbr<T>(q, ax, ax == 0 ? i0 : i1, 1, ax == 0 ? q.n0 : q.n1, d_lo, d_hi, has_lo, has_hi);
            pxb_tv_taps_col<T, VEC>(wb, ax == 0 ? q.s0 : q.s1, wc, q.cp[k], q.c0[k], q.cm[k], has_lo, has_hi, kw, d_lo, d_hi);
        }
        for (int j = 0; j < VEC; ++j) {
            zo[k][j] = zc.v[j];
            p[k][j] = zc.v[j] + q.sigma * kw[j];
        }
    }
    T a0 = T(0), a1 = T(0);
    for (int j = 0; j < VEC; ++j) {
        T grp[PXB_MAX_DIRS];
        for (int k = 0; k < NDIR; ++k) grp[k] = p[k][j];
        pxb_dual_prox_group<T>(q.hkind, NDIR, q.lam, q.sigma, grp);
        for (int k = 0; k < NDIR; ++k) {
            const T zn = q.one_m_rho * zo[k][j] + q.rho * grp[k];
            if (NORMS) {
                const T dd = zn - zo[k][j];
                a0 += dd * dd;
                a1 += zo[k][j] * zo[k][j];
            }
            p[k][j] = zn;
        }
    }
    for (int k = 0; k < NDIR; ++k) {
        PxbVec<T, VEC> o;
        for (int j = 0; j < VEC; ++j) o.v[j] = p[k][j];
        pxb_vstore<T, VEC>(zb + k * q.vol, o);
    }
    if (NORMS) { nrm[0] += (double)a0; nrm[1] += (double)a1; }
}

// ---------------------------------------------------------------------------------------------------------
// Gradient stack on its own (LinOp.apply / adjoint of a first-order Gradient): VEC voxels per thread.
//   apply  : z_k[s] = cm*x[s-e_k] + c0*x[s] + cp*x[s+e_k]          (read 1 array, write NDIR)
//   adjoint: x[s]   = sum_k cm*z_k[s+e_k] + c0*z_k[s] + cp*z_k[s-e_k]
// Faces with a folding boundary mode fall back, per voxel, to the generic bodies (as the fused half-steps do).
// ---------------------------------------------------------------------------------------------------------
template <class T, int NDIR, int VEC>
PXB_HD void pxb_tv_grad_apply_vec(const PxbTvP<T>& q, const pxb_grad_desc& d, const T* __restrict__ x, T* __restrict__ z, int64_t b, int i0,
                                  int i1, int i2) {
    const int64_t v = (int64_t)i0 * q.s0 + (int64_t)i1 * q.s1 + i2;
    const T* __restrict__ xb = x + b * q.vol + v;
    T* __restrict__ zb = z + b * NDIR * q.vol + v;
    const PxbVec<T, VEC> xc = pxb_vload<T, VEC>(xb);
    for (int k = 0; k < NDIR; ++k) {
        const int ax = 3 - NDIR + k;
        PxbVec<T, VEC> o;
        int d_lo, d_hi;
        bool has_lo, has_hi;
        if (ax == 2) {
            pxb_tv_nbr<T>(q, 2, i2, VEC, q.n2, d_lo, d_hi, has_lo, has_hi);
            pxb_tv_taps_row<T, VEC>(xb, xc, q.cp[k], q.c0[k], q.cm[k], has_lo, has_hi, o.v, d_lo, VEC - 1 + d_hi);
        } else {
            pxb_tv_nbr<T>(q, ax, ax == 0 ? i0 : i1, 1, ax == 0 ? q.n0 : q.n1, d_lo, d_hi, has_lo, has_hi);
            pxb_tv_taps_col<T, VEC>(xb, ax == 0 ? q.s0 : q.s1, xc, q.cp[k], q.c0[k], q.cm[k], has_lo, has_hi, o.v, d_lo, d_hi);
        }
        pxb_vstore<T, VEC>(zb + k * q.vol, o);
    }
}

template <class T, int NDIR, int VEC>
PXB_HD void pxb_tv_grad_adjoint_vec(const PxbTvP<T>& q, const pxb_grad_desc& d, const T* __restrict__ z, T* __restrict__ x, int64_t b, int i0,
                                    int i1, int i2) {
    const int64_t v = (int64_t)i0 * q.s0 + (int64_t)i1 * q.s1 + i2;
    const T* __restrict__ zb = z + b * NDIR * q.vol + v;
    PxbVec<T, VEC> acc;
    for (int j = 0; j < VEC; ++j) acc.v[j] = T(0);
    for (int k = 0; k < NDIR; ++k) {
        const int ax = 3 - NDIR + k;
        const T* __restrict__ zk = zb + k * q.vol;
        const PxbVec<T, VEC> c = pxb_vload<T, VEC>(zk);
        T t[VEC];
        if (ax == 2) {
            pxb_tv_taps_row<T, VEC>(zk, c, q.cm[k], q.c0[k], q.cp[k], i2 > 0, i2 + VEC < q.n2, t);
        } else {
            const int i = ax == 0 ? i0 : i1, n = ax == 0 ? q.n0 : q.n1;
            const bool has_lo = i > 0 || (ax == 0 && q.open_lo), has_hi = i < n - 1 || (ax == 0 && q.open_hi);
            pxb_tv_taps_col<T, VEC>(zk, ax == 0 ? q.s0 : q.s1, c, q.cm[k], q.c0[k], q.cp[k], has_lo, has_hi, t);
        }
        for (int j = 0; j < VEC; ++j) acc.v[j] += t[j];
    }
    pxb_tv_fold_kz<T, VEC, NDIR>(q, z + b * NDIR * q.vol, i0, i1, i2, acc.v);
    pxb_vstore<T, VEC>(x + b * q.vol + v, acc);
}
