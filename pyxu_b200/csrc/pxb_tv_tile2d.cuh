// pxb_tv_tile2d.cuh -- single-kernel PD3O / CondatVu iteration for 2-D TV problems (directions along the last two
// axes), TMA-staged tiles.
//
// Same algebra as pxb_tv_iter.cuh, but a 2-D image has no axis to march along: one CTA = one tile of TY x T2 samples
// of one image.  Thread 0 issues four (five) TMA box loads -- u, shift / grad f, z_row (+-2 rows), z_col, all with one
// vector of columns on each side, zero-filled outside the image ('constant' boundary) -- the CTA computes w on the tile
// and its one-sample rim into shared memory (phase A), then the dual update of the tile (phase C) with z (old) read
// back from the staged boxes.  50 KB of shared memory per CTA -> 4 CTAs per SM overlap each other's loads and math.
// HBM traffic (fp32): read u, shift, z0, z1 + write u, z0, z1 = 28 B/voxel (+4 B when x is written); rims re-read via L2.
// (The direct-load marching form of pxb_tv_iter.cuh reached 3.5 TB/s on 8192^2: two dependent DRAM round trips per row.)
#pragma once
#include "pxb_tv_iter.cuh"

template <class T, int VEC>
struct PxbT2Cfg {
    static constexpr int TXL = 32, TY = 16, NT = 256, T2 = TXL * VEC;
    static constexpr int BW = T2 + 2 * VEC, BR = TY + 2, BRZ = TY + 4;
    static constexpr int PADE = 128 / (int)sizeof(T);
    static constexpr int BOX = (BW * BR + PADE - 1) / PADE * PADE, BOXZ = (BW * BRZ + PADE - 1) / PADE * PADE;
    static constexpr int OFF_U = 0, OFF_S = BOX, OFF_ZC = 2 * BOX, OFF_ZR = 3 * BOX, OFF_W = 3 * BOX + BOXZ, TOTAL = 4 * BOX + BOXZ;
    static constexpr size_t SMEM = sizeof(T) * TOTAL;
    static constexpr uint32_t BYTES_BOX = BW * BR * sizeof(T), BYTES_BOXZ = BW * BRZ * sizeof(T);
};

struct PxbT2Geom {
    int n1, n2, n0;        // image rows, columns; images per batch item
    int64_t nimg;          // batch * n0
    int64_t s0, vol;       // elements between images, between components
    int ntR, ntC;
    int has_shift;         // per-voxel array staged in OFF_S (shift of the data term, or grad f for CondatVu)
    int sh_mode;           // 0: indexed like u (image index), 1: broadcast over the batch (image index modulo n0)
    int64_t nblocks;
};

// w / new primal iterate for W samples at box position (row br of the (TY+2)-row boxes, column bc)
// `fold` (MODES instances, in-image samples only): z of the batch item (component 0, image 0) for the fold terms of
// K^T z (pxb_tv_fold_kz) at sample (f0, f1, f2); null otherwise.
template <class T, int VEC, int W, int ALGO, class S, bool MODES = false>
PXB_HD void pxb_t2_w(const PxbTvP<T>& q, const PxbT2Geom& g, const T* __restrict__ sm, int br, int bc, T* wv, T* xo, T* un, T* uold,
                     const T* __restrict__ fold = nullptr, int f0 = 0, int f1 = 0, int f2 = 0) {
    using C = PxbT2Cfg<T, VEC>;
    const int i = br * C::BW + bc, iz = (br + 1) * C::BW + bc;
    T kz[W];
    {   // direction 0 acts along the rows:  (K^T z)[s] = cm z[s+e] + c0 z[s] + cp z[s-e]
        const T* __restrict__ z = sm + C::OFF_ZR + iz;
        const PxbVec<T, W> c = pxb_vload<T, W>(z);
        for (int j = 0; j < W; ++j) kz[j] = q.c0[0] * c.v[j];
        if (pxb_has_cm<S>(q, 0)) { const PxbVec<T, W> n = pxb_vload<T, W>(z + C::BW); for (int j = 0; j < W; ++j) kz[j] += q.cm[0] * n.v[j]; }
        if (pxb_has_cp<S>(q, 0)) { const PxbVec<T, W> n = pxb_vload<T, W>(z - C::BW); for (int j = 0; j < W; ++j) kz[j] += q.cp[0] * n.v[j]; }
    }
    {   // direction 1 along the row
        const T* __restrict__ z = sm + C::OFF_ZC + i;
        const PxbVec<T, W> c = pxb_vload<T, W>(z);
        T lo = T(0), hi = T(0);
        if (pxb_has_cp<S>(q, 1)) lo = z[-1];
        if (pxb_has_cm<S>(q, 1)) hi = z[W];
        for (int j = 0; j < W; ++j) {
            kz[j] += q.c0[1] * c.v[j];
            if (pxb_has_cm<S>(q, 1)) kz[j] += q.cm[1] * (j + 1 < W ? c.v[j + 1 < W ? j + 1 : 0] : hi);
            if (pxb_has_cp<S>(q, 1)) kz[j] += q.cp[1] * (j > 0 ? c.v[j > 0 ? j - 1 : 0] : lo);
        }
    }
    if (MODES && fold) pxb_tv_fold_kz<T, W, 2, S::SCHEME != PXB_SCHEME_FWD>(q, fold, f0, f1, f2, kz);
    const PxbVec<T, W> old = pxb_vload<T, W>(sm + C::OFF_U + i);
    PxbVec<T, W> sh;
    for (int j = 0; j < W; ++j) sh.v[j] = T(0);
    if (g.has_shift) sh = pxb_vload<T, W>(sm + C::OFF_S + i);
    else if (q.fkind == PXB_F_SQL2 && q.shift_mode == PXB_SHIFT_SCALAR) { for (int j = 0; j < W; ++j) sh.v[j] = q.shift[0]; }
    const int gk = pxb_gkind<S>(q);
    for (int j = 0; j < W; ++j) {
        uold[j] = old.v[j];
        if (ALGO == PXB_PD3O) {
            const T x = pxb_prox_eval<T>(gk, q.gp0, q.gp1, old.v[j] - q.tau * kz[j], q.tau);
            const T gf = (q.fkind == PXB_F_SQL2) ? (x + sh.v[j]) * q.two_alpha : T(0);
            const T ut = x - q.tau * gf;
            wv[j] = x + ut - old.v[j];
            un[j] = ut;
            xo[j] = x;
        } else {
            T gf = T(0);
            if (q.fkind == PXB_F_SQL2) gf = (old.v[j] + sh.v[j]) * q.two_alpha;
            else if (q.fkind == PXB_F_GRADARR) gf = sh.v[j];
            const T xt = pxb_prox_eval<T>(gk, q.gp0, q.gp1, old.v[j] - q.tau * gf - q.tau * kz[j], q.tau);
            wv[j] = T(2) * xt - old.v[j];
            un[j] = xt;
        }
    }
    if (!q.rho1)  // (uniform; rho == 1: the relaxed iterate IS the new one, bit for bit)
        for (int j = 0; j < W; ++j) un[j] = q.one_m_rho * old.v[j] + q.rho * un[j];
    if (ALGO != PXB_PD3O)
        for (int j = 0; j < W; ++j) xo[j] = un[j];
}

struct PxbT2Item {
    int64_t img, b;
    int i0, r0, c0;
};
PXB_HD PxbT2Item pxb_t2_item(const PxbT2Geom& g, int64_t blk, int ty, int t2) {
    PxbT2Item it;
    const int tC = (int)(blk % g.ntC); blk /= g.ntC;
    const int tR = (int)(blk % g.ntR); blk /= g.ntR;
    it.img = blk;
    it.b = blk / g.n0;
    it.i0 = (int)(blk - it.b * g.n0);
    it.r0 = tR * ty;
    it.c0 = tC * t2;
    return it;
}

// phase A for thread `tid`: w of rows {wy, wy+8} of the tile (+ the rims this thread owns) -> shared memory; new primal
// iterate (and x) of the tile's own samples -> global memory.
// MODES (folding boundary modes): as in pxb_tv_tma.cuh -- in-image samples add the fold terms of K^T z, the cells of the
// w tile one step outside the image receive w at the sample the boundary map folds them onto.
template <class T, int VEC, int ALGO, bool NORMS, class S, bool MODES = false>
PXB_HD void pxb_t2_phaseA(const PxbTvP<T>& q, const PxbT2Geom& g, const PxbT2Item& it, const PxbIterPtr<T>& a, int tid, T* __restrict__ sm, double* acc) {
    using C = PxbT2Cfg<T, VEC>;
    T* __restrict__ wsm = sm + C::OFF_W;
    const int lane = tid & 31, wy = tid >> 5, cl = lane * VEC;
    const int64_t base = it.img * g.s0;
    // MODES: the image's z for the fold terms of K^T z -- null when the tile (with its rim) holds no fold target: interior tiles
    // skip the fold code with this one uniform test (every tile paid its per-thread tests before: 8192^2 'reflect' 1.27 x 'constant')
    auto hit = [](int t, int lo, int hi) { return t != PXB_NOSRC && t >= lo && t <= hi; };
    const bool anyf = MODES && (hit(q.fold_hi[1], it.r0 - 1, it.r0 + C::TY) || hit(q.fold_lo[1], it.r0 - 1, it.r0 + C::TY) ||
                                hit(q.fold_hi[2], it.c0 - 1, it.c0 + C::T2) || hit(q.fold_lo[2], it.c0 - 1, it.c0 + C::T2));
    const T* __restrict__ zfold = anyf ? a.z_in + it.b * 2 * g.vol : nullptr;
    for (int half = 0; half < 2; ++half) {
        const int rl = wy + 8 * half, r = it.r0 + rl, c = it.c0 + cl;
        const bool in = r < g.n1 && c < g.n2;
        T wv[VEC], xo[VEC], un[VEC], uo[VEC];
        pxb_t2_w<T, VEC, VEC, ALGO, S, MODES>(q, g, sm, rl + 1, cl + VEC, wv, xo, un, uo, in ? zfold : nullptr, it.i0, r, c);
        if (MODES && !in) pxb_tv_w_outside<T, VEC, 2, ALGO>(q, a.u_in, a.z_in, it.b, it.i0, r, c, wv);
        const bool keep = MODES || in;
        PxbVec<T, VEC> o;
        for (int j = 0; j < VEC; ++j) o.v[j] = keep ? wv[j] : T(0);
        pxb_vstore<T, VEC>(wsm + (rl + 1) * C::BW + cl + VEC, o);
        if (in) {
            const int64_t lin = base + (int64_t)r * g.n2 + c;
            if (ALGO == PXB_PD3O) {
                if (NORMS && a.norms_x) {
                    const PxbVec<T, VEC> xp = pxb_vload<T, VEC>(a.x_out + lin);
                    T s0 = T(0), s1 = T(0);  // one vector's partial sums in the working precision, widened once
                    for (int j = 0; j < VEC; ++j) {
                        const T dd = xo[j] - xp.v[j];
                        s0 += dd * dd;
                        s1 += xp.v[j] * xp.v[j];
                    }
                    acc[0] += (double)s0;
                    acc[1] += (double)s1;
                }
                if (a.x_out) { for (int j = 0; j < VEC; ++j) o.v[j] = xo[j]; pxb_vstore<T, VEC>(a.x_out + lin, o); }
            } else if (NORMS && a.norms_x) {
                T s0 = T(0), s1 = T(0);
                for (int j = 0; j < VEC; ++j) {
                    const T dd = un[j] - uo[j];
                    s0 += dd * dd;
                    s1 += uo[j] * uo[j];
                }
                acc[0] += (double)s0;
                acc[1] += (double)s1;
            }
            for (int j = 0; j < VEC; ++j) o.v[j] = un[j];
            pxb_vstore<T, VEC>(a.u_out + lin, o);
        }
    }
    if (wy < 2) {  // rim rows: r0-1 (needed when cm != 0 along the rows), r0+TY (cp != 0)
        const bool top = wy == 0;
        if (top ? pxb_has_cm<S>(q, 0) : pxb_has_cp<S>(q, 0)) {
            const int br = top ? 0 : C::TY + 1, r = top ? it.r0 - 1 : it.r0 + C::TY, c = it.c0 + cl;
            T wv[VEC], xo[VEC], un[VEC], uo[VEC];
            PxbVec<T, VEC> o;
            if (MODES) {  // the row this rim row stands for (pxb_rim_src), evaluated from the tile's own boxes when it lies in the tile
                const int rs = pxb_rim_src(r, g.n1, q.mode[1], it.r0, C::TY, true);
                if (rs != PXB_NOSRC && c < g.n2)
                    pxb_t2_w<T, VEC, VEC, ALGO, S, true>(q, g, sm, rs - (it.r0 - 1), cl + VEC, wv, xo, un, uo, zfold, it.i0, rs, c);
                else pxb_tv_w_outside<T, VEC, 2, ALGO>(q, a.u_in, a.z_in, it.b, it.i0, r, c, wv);  // a fold onto another tile ('wrap'), or zeros
                for (int j = 0; j < VEC; ++j) o.v[j] = wv[j];
            } else {
                pxb_t2_w<T, VEC, VEC, ALGO, S>(q, g, sm, br, cl + VEC, wv, xo, un, uo);
                const bool in = r >= 0 && r < g.n1 && c < g.n2;
                for (int j = 0; j < VEC; ++j) o.v[j] = in ? wv[j] : T(0);
            }
            pxb_vstore<T, VEC>(wsm + br * C::BW + cl + VEC, o);
        }
    }
    if (wy == 7) {  // rim columns: lanes 0..15 the left one (c0-1), lanes 16..31 the right one (c0+T2)
        const bool left = lane < C::TY;
        const int rl = left ? lane : lane - C::TY;
        if (left ? pxb_has_cm<S>(q, 1) : pxb_has_cp<S>(q, 1)) {
            const int bc = left ? VEC - 1 : VEC + C::T2, r = it.r0 + rl, c = left ? it.c0 - 1 : it.c0 + C::T2;
            T wv[1], xo[1], un[1], uo[1];
            if (MODES) {
                const int cs = pxb_rim_src(c, g.n2, q.mode[2], it.c0, C::T2, true);
                if (r < g.n1 && cs != PXB_NOSRC)
                    pxb_t2_w<T, VEC, 1, ALGO, S, true>(q, g, sm, rl + 1, cs - (it.c0 - VEC), wv, xo, un, uo, zfold, it.i0, r, cs);
                else pxb_tv_w_outside<T, 1, 2, ALGO>(q, a.u_in, a.z_in, it.b, it.i0, r, c, wv);
                wsm[(rl + 1) * C::BW + bc] = wv[0];
            } else {
                pxb_t2_w<T, VEC, 1, ALGO, S>(q, g, sm, rl + 1, bc, wv, xo, un, uo);
                wsm[(rl + 1) * C::BW + bc] = (r < g.n1 && c >= 0 && c < g.n2) ? wv[0] : T(0);
            }
        }
    }
}

// phase C: z_out = (1-rho) z + rho prox_{sigma h*}(z + sigma K w) on the tile (z from the staged boxes, w from shared memory)
template <class T, int VEC, bool NORMS, class S>
PXB_HD void pxb_t2_phaseC(const PxbTvP<T>& q, const PxbT2Geom& g, const PxbT2Item& it, const PxbIterPtr<T>& a, int tid, const T* __restrict__ sm,
                          double* acc) {
    using C = PxbT2Cfg<T, VEC>;
    const T* __restrict__ wsm = sm + C::OFF_W;
    const int lane = tid & 31, wy = tid >> 5, cl = lane * VEC;
    for (int half = 0; half < 2; ++half) {
        const int rl = wy + 8 * half, r = it.r0 + rl, c = it.c0 + cl;
        if (r >= g.n1 || c >= g.n2) continue;
        const int cell = (rl + 1) * C::BW + cl + VEC;
        const T* __restrict__ s1 = wsm + cell;
        const PxbVec<T, VEC> wc = pxb_vload<T, VEC>(s1);
        const PxbVec<T, VEC> z0 = pxb_vload<T, VEC>(sm + C::OFF_ZR + cell + C::BW), z1 = pxb_vload<T, VEC>(sm + C::OFF_ZC + cell);
        T p[2][VEC];
        {   // (K w)[s] = cm w[s-e] + c0 w[s] + cp w[s+e] along the rows
            T kw[VEC];
            for (int j = 0; j < VEC; ++j) kw[j] = q.c0[0] * wc.v[j];
            if (pxb_has_cm<S>(q, 0)) { const PxbVec<T, VEC> n = pxb_vload<T, VEC>(s1 - C::BW); for (int j = 0; j < VEC; ++j) kw[j] += q.cm[0] * n.v[j]; }
            if (pxb_has_cp<S>(q, 0)) { const PxbVec<T, VEC> n = pxb_vload<T, VEC>(s1 + C::BW); for (int j = 0; j < VEC; ++j) kw[j] += q.cp[0] * n.v[j]; }
            for (int j = 0; j < VEC; ++j) p[0][j] = z0.v[j] + q.sigma * kw[j];
        }
        {
            T lo = T(0), hi = T(0);
            if (pxb_has_cm<S>(q, 1)) lo = s1[-1];
            if (pxb_has_cp<S>(q, 1)) hi = s1[VEC];
            for (int j = 0; j < VEC; ++j) {
                T kw = q.c0[1] * wc.v[j];
                if (pxb_has_cp<S>(q, 1)) kw += q.cp[1] * (j + 1 < VEC ? wc.v[j + 1 < VEC ? j + 1 : 0] : hi);
                if (pxb_has_cm<S>(q, 1)) kw += q.cm[1] * (j > 0 ? wc.v[j > 0 ? j - 1 : 0] : lo);
                p[1][j] = z1.v[j] + q.sigma * kw;
            }
        }
        PxbVec<T, VEC> o0, o1;
        T a0 = T(0), a1 = T(0);
        for (int j = 0; j < VEC; ++j) {
            T grp[PXB_MAX_DIRS] = {p[0][j], p[1][j], T(0)};
            pxb_dual_prox_group<T>(pxb_hkind<S>(q), 2, q.lam, q.sigma, grp);
            o0.v[j] = grp[0];
            o1.v[j] = grp[1];
        }
        if (!q.rho1)  // (uniform; rho == 1: (1 - rho) z + rho p is p, bit for bit)
            for (int j = 0; j < VEC; ++j) {
                o0.v[j] = q.one_m_rho * z0.v[j] + q.rho * o0.v[j];
                o1.v[j] = q.one_m_rho * z1.v[j] + q.rho * o1.v[j];
            }
        if (NORMS)
            for (int j = 0; j < VEC; ++j) {
                const T d0 = o0.v[j] - z0.v[j], d1 = o1.v[j] - z1.v[j];
                a0 += d0 * d0 + d1 * d1;
                a1 += z0.v[j] * z0.v[j] + z1.v[j] * z1.v[j];
            }
        if (NORMS) { acc[2] += (double)a0; acc[3] += (double)a1; }
        T* __restrict__ zb = a.z_out + it.b * 2 * g.vol + (int64_t)it.i0 * g.s0 + (int64_t)r * g.n2 + c;
        pxb_vstore<T, VEC>(zb, o0);
        pxb_vstore<T, VEC>(zb + g.vol, o1);
    }
}

// host: eligibility + geometry.  0 or a reason code.
template <class T, int VEC>
inline int pxb_t2_setup(const pxb_grad_desc& d, const pxb_pds_params& P, PxbTvCoef& cf, PxbTvP<T>& q, PxbT2Geom& g, bool allow_modes = false) {
    using C = PxbT2Cfg<T, VEC>;
    if (!pxb_tv_fast_coefs(d, cf)) return 1;
    if (d.ndir != 2) return 2;
    if (P.hkind != PXB_DUAL_L21 && P.hkind != PXB_DUAL_L1) return 3;
    if (pxb_any_mode(d) && !allow_modes) return 4;
    if (d.shape[2] % VEC) return 5;
    if (d.slab.halo != 0 || d.slab.open_lo || d.slab.open_hi) return 7;
    if (d.shape[0] < 1 || d.shape[1] < 1 || d.shape[2] < 1 || d.batch < 1) return 9;
    pxb_tv_prepare<T>(d, cf, P, q);
    g.n0 = (int)d.shape[0]; g.n1 = (int)d.shape[1]; g.n2 = (int)d.shape[2];
    g.nimg = d.batch * g.n0;
    g.s0 = (int64_t)g.n1 * g.n2;
    g.vol = g.s0 * g.n0;
    g.ntR = (g.n1 + C::TY - 1) / C::TY;
    g.ntC = (g.n2 + C::T2 - 1) / C::T2;
    g.nblocks = (int64_t)g.ntR * g.ntC * g.nimg;
    if (g.nblocks > 0x7fffffffLL) return 8;
    g.has_shift = 0;
    g.sh_mode = 0;
    if (q.fkind == PXB_F_GRADARR) g.has_shift = 1;
    else if (q.fkind == PXB_F_SQL2) {
        if (q.shift_mode == PXB_SHIFT_LIN) g.has_shift = 1;
        else if (q.shift_mode == PXB_SHIFT_VOL) { g.has_shift = 1; g.sh_mode = 1; }
        else if (q.shift_mode == PXB_SHIFT_MOD) return 21;
    }
    return 0;
}
