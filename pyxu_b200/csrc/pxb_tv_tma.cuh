// pxb_tv_tma.cuh -- TMA-staged form of the single-kernel PD3O / CondatVu iteration for 3-D volumes.
//
// Same algorithm, tiles, shared-memory w ring and phase C as pxb_tv_iter.cuh, but phase A no longer issues global
// loads: one thread per CTA asks the Tensor Memory Accelerator for the boxes of plane m+3
//        u, shift (or grad f), z0, z2 : rows [r0-1, r0+TY+1) x columns [c0-VEC, c0+T2+VEC)
//        z1                           : rows [r0-2, r0+TY+2) x the same columns
// while the CTA computes plane m out of shared memory (3-stage ring, one mbarrier per stage, completion by
// transaction bytes).  Out-of-domain rows / columns / planes are ZERO-FILLED by the TMA unit, which is exactly the
// reference's 'constant' boundary, so phase A carries no boundary branches for its operands; only w itself is masked
// outside the domain.  DRAM latency is hidden by the ring instead of by occupancy (the direct-load form exposed two
// dependent DRAM round trips per plane: 46 % of the HBM roofline).
//
// The arithmetic below is __host__ __device__ and is replayed on the CPU by tests/emu with the box loads emulated
// by a plain gather with zero fill (pxb_tma_box_desc is what both the tensor-map encoder and the emulation consume).
#pragma once
#include "pxb_tv_iter.cuh"

template <class T, int VEC, int TY>
struct PxbTmaCfg {
    static constexpr int TXL = 32, NT = TXL * TY, T2 = TXL * VEC;
    static constexpr int BW = T2 + 2 * VEC;         // box width (elements): one vector of rim on each side
    static constexpr int BR = TY + 2, BR1 = TY + 4;  // box rows: +-1 (u, shift, z0, z2), +-2 (z1)
    static constexpr int PADE = 128 / (int)sizeof(T);
    static constexpr int BOX = (BW * BR + PADE - 1) / PADE * PADE;    // elements, 128-byte multiples (TMA destination alignment)
    static constexpr int BOX1 = (BW * BR1 + PADE - 1) / PADE * PADE;
    static constexpr int OFF_U = 0, OFF_S = BOX, OFF_Z0 = 2 * BOX, OFF_Z2 = 3 * BOX, OFF_Z1 = 4 * BOX, STAGE = 4 * BOX + BOX1;
    static constexpr int NSTAGE = 3;
    using Ring = PxbIterCfg<T, VEC, TXL, TY, 3>;     // w ring: identical layout to the direct-load form (RS == BW)
    static constexpr size_t SMEM_STAGES = sizeof(T) * NSTAGE * STAGE, SMEM_RING = Ring::SMEM;
    static constexpr size_t SMEM = SMEM_STAGES + SMEM_RING + 64;      // + mbarriers
    static constexpr uint32_t BYTES_BOX = BW * BR * sizeof(T), BYTES_BOX1 = BW * BR1 * sizeof(T);
};

// One tensor map = one 5-D view (columns, rows, planes, component, batch) of an array.
struct PxbTmaBoxDesc {
    const void* base;
    uint64_t dim[5];
    uint64_t stride[5];  // elements; stride[0] == 1
    uint32_t box[5];
};

struct PxbTmaGeom {
    int gl;         // ghost planes below owned plane 0 that the maps cover (plane coordinate = m + gl)
    int has_shift;  // 1: a per-voxel array (shift of the data term, or grad f for CondatVu) is staged in OFF_S
    int sh_batched; // 1: that array has one volume per batch item, 0: one volume broadcast over the batch
};

template <class T, int VEC, int TY>
inline int pxb_tma_setup(const pxb_grad_desc& d, const pxb_pds_params& P, const PxbTvCoef& cf, const PxbIterGeom& g, const PxbTvP<T>& q,
                         const void* u_in, const void* z_in, PxbTmaGeom& tg, PxbTmaBoxDesc& mu, PxbTmaBoxDesc& ms, PxbTmaBoxDesc& mz) {
    using C = PxbTmaCfg<T, VEC, TY>;
    if (g.ndir != 3) return 20;
    const bool two_sided = cf.cm[0] != 0.0 && cf.cp[0] != 0.0;
    const int gdepth = two_sided ? 2 : 1;
    tg.gl = g.open_lo ? gdepth : 0;
    const int gh = g.open_hi ? gdepth : 0;
    const uint64_t planes = (uint64_t)(g.nM + tg.gl + gh);
    auto fill = [&](PxbTmaBoxDesc& m, const void* base, uint64_t ncomp, uint64_t comp_stride, uint64_t nbatch, uint64_t batch_stride, int rows) {
        m.base = (const void*)((const T*)base - (int64_t)tg.gl * g.sM);
        m.dim[0] = g.nC; m.dim[1] = g.nR; m.dim[2] = planes; m.dim[3] = ncomp; m.dim[4] = nbatch;
        m.stride[0] = 1; m.stride[1] = g.sR; m.stride[2] = g.sM; m.stride[3] = comp_stride; m.stride[4] = batch_stride;
        m.box[0] = C::BW; m.box[1] = rows; m.box[2] = m.box[3] = m.box[4] = 1;
    };
    const uint64_t batch = (uint64_t)g.nimg;
    fill(mu, u_in, 1, g.vol, batch, g.vol, C::BR);
    fill(mz, z_in, 3, g.vol, batch, 3 * (uint64_t)g.vol, C::BR);  // box rows: BR for z0 / z2, BR1 for z1 -> z1 has its own map copy
    tg.has_shift = 0;
    tg.sh_batched = 0;
    ms = mu;
    if (q.fkind == PXB_F_GRADARR) {
        tg.has_shift = 1; tg.sh_batched = 1;
        fill(ms, q.garr, 1, g.vol, batch, g.vol, C::BR);
    } else if (q.fkind == PXB_F_SQL2) {
        if (q.shift_mode == PXB_SHIFT_LIN) { tg.has_shift = 1; tg.sh_batched = 1; fill(ms, q.shift, 1, g.vol, batch, g.vol, C::BR); }
        else if (q.shift_mode == PXB_SHIFT_VOL) { tg.has_shift = 1; fill(ms, q.shift, 1, g.vol, 1, g.vol, C::BR); }
        else if (q.shift_mode == PXB_SHIFT_MOD) return 21;
    }
    // TMA limits: strides in bytes multiples of 16 (rows are, since nC % VEC == 0), boxes <= 256 per dimension
    if ((g.sR * sizeof(T)) % 16 || (g.vol * sizeof(T)) % 16) return 22;
    return 0;
}

// which compiled instance serves a problem: 1 / 2 = forward differences + L21 + per-voxel shifted squared-l2 data
// term with g = positivity / none; 3 / 4 = the same with grad f handed over as an array (CondatVu); 0 = the generic instance
using PxbSpecFwdPos = PxbSpec<PXB_SCHEME_FWD, PXB_PROX_POS, PXB_DUAL_L21, 1>;
using PxbSpecFwdNone = PxbSpec<PXB_SCHEME_FWD, PXB_PROX_NONE, PXB_DUAL_L21, 1>;
using PxbSpecFwdPosG = PxbSpec<PXB_SCHEME_FWD, PXB_PROX_POS, PXB_DUAL_L21, 2>;   // CondatVu, grad f array
using PxbSpecFwdNoneG = PxbSpec<PXB_SCHEME_FWD, PXB_PROX_NONE, PXB_DUAL_L21, 2>;
template <class T>
inline int pxb_tma_pick_spec(const PxbTvCoef& cf, const PxbTvP<T>& q, const PxbTmaGeom& tg) {
    bool fwd = true;
    for (int k = 0; k < 3; ++k) fwd = fwd && cf.cm[k] == 0.0 && cf.cp[k] != 0.0;
    if (!(fwd && q.hkind == PXB_DUAL_L21 && tg.has_shift)) return 0;
    const int g = q.gkind == PXB_PROX_POS ? 1 : (q.gkind == PXB_PROX_NONE ? 2 : 0);
    if (g == 0) return 0;
    if (q.fkind == PXB_F_SQL2) return g;           // 1, 2
    if (q.fkind == PXB_F_GRADARR) return 2 + g;    // 3, 4 (CondatVu only)
    return 0;
}

// MODES: where the sources of the in-plane fold terms of K^T z (pxb_tv_fold_kz) sit in a work item's staged boxes.  For reflect /
// symmetric / edge the target sample (n-2 or n-1, resp. 1 or 0) and its source (n-1, resp. 0) lie in the same edge tile, so the term
// comes out of shared memory; PXB_NOSRC = not in the boxes ('wrap': the other end of the line) -> global memory.  `any_rc`: the tile
// (with its rim) holds an in-plane fold target at all -- interior tiles skip the fold code with one uniform test per plane.
// (ncu on 512^3 'reflect': the scalar global load of z[n-1] by the rim-column threads sat in the per-plane critical path -- barrier
//  stalls 3x those of the 'constant' instance, DRAM throughput 58 % against 81 %.)
struct PxbTmaFold {
    int any_rc;
    int bc_hi2, bc_lo2;  // column in the boxes of z2[.., n2-1] / z2[.., 0]
    int br_hi1, br_lo1;  // row in the z1 box of z1[n1-1, ..] / z1[0, ..]
};
template <class T, int VEC, int TY>
PXB_HD PxbTmaFold pxb_tma_fold_setup(const PxbTvP<T>& q, const PxbIterGeom& g, const PxbIterItem& it) {
    using C = PxbTmaCfg<T, VEC, TY>;
    PxbTmaFold f;
    auto hit = [](int t, int lo, int hi) { return t != PXB_NOSRC && t >= lo && t <= hi; };
    f.any_rc = (hit(q.fold_hi[1], it.r0 - 1, it.r0 + TY) || hit(q.fold_lo[1], it.r0 - 1, it.r0 + TY) || hit(q.fold_hi[2], it.c0 - 1, it.c0 + C::T2) ||
                hit(q.fold_lo[2], it.c0 - 1, it.c0 + C::T2)) ? 1 : 0;
    auto pos = [](int src, int start, int extent) { const int b = src - start; return (b >= 0 && b < extent) ? b : PXB_NOSRC; };
    f.bc_hi2 = pos(g.nC - 1, it.c0 - VEC, C::BW);
    f.bc_lo2 = pos(0, it.c0 - VEC, C::BW);
    f.br_hi1 = pos(g.nR - 1, it.r0 - 2, C::BR1);
    f.br_lo1 = pos(0, it.r0 - 2, C::BR1);
    return f;
}
// pxb_tv_fold_kz for W samples at box position (row br, column bc) = sample (f0, f1, f2): same terms, the in-plane ones out of the
// staged boxes where `fb` says they are there.
template <class T, int VEC, int TY, int W, bool LO>
PXB_HD void pxb_tma_fold_kz(const PxbTvP<T>& q, const PxbTmaFold& fb, const T* __restrict__ st, const T* __restrict__ zimg, int br, int bc, int f0, int f1,
                            int f2, T* kz) {
    using C = PxbTmaCfg<T, VEC, TY>;
    {   // along the row (component 2)
        const int th = q.fold_hi[2], tl = LO ? q.fold_lo[2] : PXB_NOSRC;
        const bool hi = th >= f2 && th < f2 + W, lo = LO && tl >= f2 && tl < f2 + W;
        if (hi || lo) {
            const T* __restrict__ row = zimg + 2 * q.vol + (int64_t)f0 * q.s0 + (int64_t)f1 * q.s1;
            if (hi) {
                const T zf = fb.bc_hi2 != PXB_NOSRC ? st[C::OFF_Z2 + br * C::BW + fb.bc_hi2] : row[q.n2 - 1];
                for (int j = 0; j < W; ++j)
                    if (f2 + j == th) kz[j] += q.cp[2] * zf;
            }
            if (lo) {
                const T zf = fb.bc_lo2 != PXB_NOSRC ? st[C::OFF_Z2 + br * C::BW + fb.bc_lo2] : row[0];
                for (int j = 0; j < W; ++j)
                    if (f2 + j == tl) kz[j] += q.cm[2] * zf;
            }
        }
    }
    {   // along the rows (component 1; its box starts one row earlier)
        const bool hi = f1 == q.fold_hi[1], lo = LO && f1 == q.fold_lo[1];
        if (hi) {
            const PxbVec<T, W> f = fb.br_hi1 != PXB_NOSRC ? pxb_vload<T, W>(st + C::OFF_Z1 + fb.br_hi1 * C::BW + bc)
                                                          : pxb_vload<T, W>(zimg + q.vol + (int64_t)f0 * q.s0 + (int64_t)(q.n1 - 1) * q.s1 + f2);
            for (int j = 0; j < W; ++j) kz[j] += q.cp[1] * f.v[j];
        }
        if (lo) {
            const PxbVec<T, W> f = fb.br_lo1 != PXB_NOSRC ? pxb_vload<T, W>(st + C::OFF_Z1 + fb.br_lo1 * C::BW + bc)
                                                          : pxb_vload<T, W>(zimg + q.vol + (int64_t)f0 * q.s0 + f2);
            for (int j = 0; j < W; ++j) kz[j] += q.cm[1] * f.v[j];
        }
    }
    {   // along the planes (component 0): two planes of the volume, from global memory
        const bool hi = f0 == q.fold_hi[0], lo = LO && f0 == q.fold_lo[0];
        if (hi || lo) {
            const T* __restrict__ line = zimg + (int64_t)f1 * q.s1 + f2;
            if (hi) {
                const PxbVec<T, W> f = pxb_vload<T, W>(line + (int64_t)(q.n0 - 1) * q.s0);
                for (int j = 0; j < W; ++j) kz[j] += q.cp[0] * f.v[j];
            }
            if (lo) {
                const PxbVec<T, W> f = pxb_vload<T, W>(line);
                for (int j = 0; j < W; ++j) kz[j] += q.cm[0] * f.v[j];
            }
        }
    }
}

// A value loaded from a staged box must have ARRIVED before the thread passes the per-plane barrier: behind it thread 0 re-arms the
// stage and the TMA unit overwrites the box.  bar.sync orders the load's issue, not its completion; a load whose result is first used
// in the next plane (the carried z0 of the rims) can sit in the memory-instruction queue behind this thread's global stores -- the
// kernel is memory-bound -- long enough to read the NEXT contents of the box: one wrong rim cell in ~10^6 thread-block-planes at 1024^3,
// different from run to run (tools/check_kernel_determinism.py).  A dependent instruction makes the scoreboard wait (a self-move is
// dropped by ptxas: SASS showed the LDS directly in front of BAR.SYNC).
template <class T>
PXB_HD T pxb_landed(const PxbTvP<T>& q, T v) {
    return v * q.one;  // (q.one == 1 comes from the constant bank: the multiplication stays, and it needs v in its register)
}

template <class T, int VEC>
struct PxbTmaThread {
    T zc[3][VEC];     // z_in at this thread's own samples, plane just visited
    T zprev[3][VEC];  // plane before (phase C with lag 1)
    T z0p[VEC];       // z0 of the previous plane at: own samples ...
    T z0p_rim[VEC];   //   ... the rim-row samples this thread computes (warps 0, 1)
    T z0p_col;        //   ... the rim-column sample (last 2*TY threads)
    T xp[VEC];        // PD3O + RelError[x]: the previous x of the plane phase A visits next, loaded one plane ahead so that
                      // the global-load latency is not in the per-plane critical path (everything else is staged by TMA)
    int64_t lin;      // offset of this thread's own samples on the plane phase A visits next, in u-like arrays (running: += sM per plane)
    T* pz;            // z_out (component 0) at this thread's own samples on the plane phase C finishes next (running)
    double acc[4];
};

// w / new primal iterate for W samples whose box position is (row br, column bc) -- everything from shared memory.
// `fold` (MODES instances, in-domain samples only): the image's z (component 0, sample (0, 0, 0)) for the fold terms of
// K^T z (pxb_tv_fold_kz) at sample (f0, f1, f2); null otherwise.
template <class T, int VEC, int TY, int W, int ALGO, class S, bool MODES = false>
PXB_HD void pxb_tma_w(const PxbTvP<T>& q, const PxbTmaGeom& tg, const T* __restrict__ st, const T* __restrict__ st_next, int br, int bc,
                      const T* z0p, T* wv, T* z0c, T* z1c, T* z2c, T* xo, T* un, T* uold, const T* __restrict__ fold = nullptr, int f0 = 0,
                      int f1 = 0, int f2 = 0, const PxbTmaFold* fb = nullptr) {
    using C = PxbTmaCfg<T, VEC, TY>;
    const int i = br * C::BW + bc, i1 = (br + 1) * C::BW + bc;
#if defined(__CUDA_ARCH__)
    // Packed form for the headline instances (fp32 vectors of the tile's own samples, PD3O, three forward differences, per-voxel
    // shifted squared-l2 data term, g = positivity | none, rho == 1): the same arithmetic on pairs of samples (fma / mul / add.f32x2),
    // -tau folded into the taps of K^T z, and x - tau grad f(x) = (1 - 2 alpha tau) x - 2 alpha tau shift.  See pxb_iter_phaseC_f32x2.
    if constexpr (sizeof(T) == 4 && W == 4 && VEC == 4 && ALGO == PXB_PD3O && S::SCHEME == PXB_SCHEME_FWD && S::FK == 1 &&
                  (S::GK == PXB_PROX_POS || S::GK == PXB_PROX_NONE) && !MODES) {
        if (q.rho1 && !PXB_EXP(64)) {  // (uniform)
            const float4 c0v = *reinterpret_cast<const float4*>(st + C::OFF_Z0 + i);
            const float4 c1v = *reinterpret_cast<const float4*>(st + C::OFF_Z1 + i1);
            const float4 n1v = *reinterpret_cast<const float4*>(st + C::OFF_Z1 + i1 - C::BW);
            const float4 c2v = *reinterpret_cast<const float4*>(st + C::OFF_Z2 + i);
            const float lo = st[C::OFF_Z2 + i - 1];
            const float4 old = *reinterpret_cast<const float4*>(st + C::OFF_U + i);
            const float4 shv = *reinterpret_cast<const float4*>(st + C::OFF_S + i);
            auto pk = [&](int i) { return *reinterpret_cast<const float2*>(q.pk[i]); };  // (c, c) pairs folded on the host
            const float2 t00 = pk(6), t10 = pk(7), t20 = pk(8), t0p = pk(9), t1p = pk(10), t2p = pk(11), a1 = pk(12), a2 = pk(13), m1 = pk(14);
            float2 vl = __ffma2_rn(t00, make_float2(c0v.x, c0v.y), make_float2(old.x, old.y));
            float2 vh = __ffma2_rn(t00, make_float2(c0v.z, c0v.w), make_float2(old.z, old.w));
            vl = __ffma2_rn(t0p, make_float2(z0p[0], z0p[1]), vl);
            vh = __ffma2_rn(t0p, make_float2(z0p[2], z0p[3]), vh);
            vl = __ffma2_rn(t10, make_float2(c1v.x, c1v.y), vl);
            vh = __ffma2_rn(t10, make_float2(c1v.z, c1v.w), vh);
            vl = __ffma2_rn(t1p, make_float2(n1v.x, n1v.y), vl);
            vh = __ffma2_rn(t1p, make_float2(n1v.z, n1v.w), vh);
            vl = __ffma2_rn(t20, make_float2(c2v.x, c2v.y), vl);
            vh = __ffma2_rn(t20, make_float2(c2v.z, c2v.w), vh);
            vl = __ffma2_rn(t2p, make_float2(lo, c2v.x), vl);
            vh = __ffma2_rn(t2p, make_float2(c2v.y, c2v.z), vh);
            float2 xl = vl, xh = vh;
            if (S::GK == PXB_PROX_POS) { xl = make_float2(fmaxf(vl.x, 0.f), fmaxf(vl.y, 0.f)); xh = make_float2(fmaxf(vh.x, 0.f), fmaxf(vh.y, 0.f)); }
            const float2 ul = __ffma2_rn(a1, xl, __fmul2_rn(a2, make_float2(shv.x, shv.y)));
            const float2 uh = __ffma2_rn(a1, xh, __fmul2_rn(a2, make_float2(shv.z, shv.w)));
            const float2 wl = __ffma2_rn(m1, make_float2(old.x, old.y), __fadd2_rn(xl, ul));
            const float2 wh = __ffma2_rn(m1, make_float2(old.z, old.w), __fadd2_rn(xh, uh));
            wv[0] = wl.x; wv[1] = wl.y; wv[2] = wh.x; wv[3] = wh.y;
            un[0] = ul.x; un[1] = ul.y; un[2] = uh.x; un[3] = uh.y;
            xo[0] = xl.x; xo[1] = xl.y; xo[2] = xh.x; xo[3] = xh.y;
            uold[0] = old.x; uold[1] = old.y; uold[2] = old.z; uold[3] = old.w;
            z0c[0] = c0v.x; z0c[1] = c0v.y; z0c[2] = c0v.z; z0c[3] = c0v.w;
            z1c[0] = c1v.x; z1c[1] = c1v.y; z1c[2] = c1v.z; z1c[3] = c1v.w;
            z2c[0] = c2v.x; z2c[1] = c2v.y; z2c[2] = c2v.z; z2c[3] = c2v.w;
            return;
        }
    }
#endif
    T kz[W];
    {   // along M: (K^T z)[s] = cm z[s+e] + c0 z[s] + cp z[s-e]
        const PxbVec<T, W> c = pxb_vload<T, W>(st + C::OFF_Z0 + i);
        for (int j = 0; j < W; ++j) { z0c[j] = c.v[j]; kz[j] = q.c0[0] * c.v[j]; }
        if (pxb_has_cp<S>(q, 0)) for (int j = 0; j < W; ++j) kz[j] += q.cp[0] * z0p[j];
        if (pxb_has_cm<S>(q, 0)) {
            const PxbVec<T, W> n = pxb_vload<T, W>(st_next + C::OFF_Z0 + i);
            for (int j = 0; j < W; ++j) kz[j] += q.cm[0] * n.v[j];
        }
    }
    {   // along the rows (the z1 box starts one row earlier); zero fill stands in for every boundary test
        const T* __restrict__ z1 = st + C::OFF_Z1 + i1;
        const PxbVec<T, W> c = pxb_vload<T, W>(z1);
        for (int j = 0; j < W; ++j) { z1c[j] = c.v[j]; kz[j] += q.c0[1] * c.v[j]; }
        if (pxb_has_cm<S>(q, 1)) {
            const PxbVec<T, W> n = pxb_vload<T, W>(z1 + C::BW);
            for (int j = 0; j < W; ++j) kz[j] += q.cm[1] * n.v[j];
        }
        if (pxb_has_cp<S>(q, 1)) {
            const PxbVec<T, W> n = pxb_vload<T, W>(z1 - C::BW);
            for (int j = 0; j < W; ++j) kz[j] += q.cp[1] * n.v[j];
        }
    }
    {   // along the row
        const T* __restrict__ z2 = st + C::OFF_Z2 + i;
        const PxbVec<T, W> c = pxb_vload<T, W>(z2);
        T lo = T(0), hi = T(0);
        if (pxb_has_cp<S>(q, 2)) lo = z2[-1];
        if (pxb_has_cm<S>(q, 2)) hi = z2[W];
        for (int j = 0; j < W; ++j) {
            z2c[j] = c.v[j];
            kz[j] += q.c0[2] * c.v[j];
            if (pxb_has_cm<S>(q, 2)) kz[j] += q.cm[2] * (j + 1 < W ? c.v[j + 1 < W ? j + 1 : 0] : hi);
            if (pxb_has_cp<S>(q, 2)) kz[j] += q.cp[2] * (j > 0 ? c.v[j > 0 ? j - 1 : 0] : lo);
        }
    }
    if (MODES && fold) pxb_tma_fold_kz<T, VEC, TY, W, S::SCHEME != PXB_SCHEME_FWD>(q, *fb, st, fold, br, bc, f0, f1, f2, kz);
    const PxbVec<T, W> old = pxb_vload<T, W>(st + C::OFF_U + i);
    PxbVec<T, W> sh;
    for (int j = 0; j < W; ++j) sh.v[j] = T(0);
    if (S::FK >= 1 || tg.has_shift) sh = pxb_vload<T, W>(st + C::OFF_S + i);
    else if (q.fkind == PXB_F_SQL2 && q.shift_mode == PXB_SHIFT_SCALAR) { for (int j = 0; j < W; ++j) sh.v[j] = q.shift[0]; }
    const int gk = pxb_gkind<S>(q);
    for (int j = 0; j < W; ++j) {
        uold[j] = old.v[j];
        if (ALGO == PXB_PD3O) {
            const T x = pxb_prox_eval<T>(gk, q.gp0, q.gp1, old.v[j] - q.tau * kz[j], q.tau);
            const T gf = (S::FK == 1 || q.fkind == PXB_F_SQL2) ? (x + sh.v[j]) * q.two_alpha : T(0);
            const T ut = x - q.tau * gf;
            wv[j] = x + ut - old.v[j];
            un[j] = ut;
            xo[j] = x;
        } else {
            T gf = T(0);
            if (S::FK == 2) gf = sh.v[j];
            else if (S::FK == 1 || q.fkind == PXB_F_SQL2) gf = (old.v[j] + sh.v[j]) * q.two_alpha;
            else if (q.fkind == PXB_F_GRADARR) gf = sh.v[j];
            const T vv = old.v[j] - q.tau * gf - q.tau * kz[j];
            const T xt = pxb_prox_eval<T>(gk, q.gp0, q.gp1, vv, q.tau);
            wv[j] = T(2) * xt - old.v[j];
            un[j] = xt;
        }
    }
    if (!q.rho1)  // (uniform; rho == 1: the relaxed iterate IS the new one, bit for bit)
        for (int j = 0; j < W; ++j) un[j] = q.one_m_rho * old.v[j] + q.rho * un[j];
    if (ALGO != PXB_PD3O)
        for (int j = 0; j < W; ++j) xo[j] = un[j];
}

// z0 of plane mlo-1 at the samples whose previous-plane value this thread carries (start of a work item)
// (MODES: a rim thread carries the value at the sample its rim cell stands for, see pxb_rim_src)
// XPREV (PD3O with RelError[x] sums): also the previous x of the first plane the work item updates
template <class T, int VEC, int TY, bool MODES = false, bool XPREV = false>
PXB_HD void pxb_tma_prologue(const PxbTvP<T>& q, const PxbIterGeom& g, const PxbIterItem& it, const PxbIterPtr<T>& a, int tid, int mlo,
                             PxbTmaThread<T, VEC>& st) {
    using C = PxbTmaCfg<T, VEC, TY>;
    for (int j = 0; j < VEC; ++j) st.z0p[j] = st.z0p_rim[j] = T(0);
    st.z0p_col = T(0);
    {   // running addresses of the thread's own samples (the per-plane 64-bit index arithmetic was ~30 of 463 warp instructions per plane)
        const int rl0 = tid / C::TXL, r_ = it.r0 + rl0, c_ = it.c0 + (tid - rl0 * C::TXL) * VEC;
        const int64_t rc = (int64_t)r_ * g.sR + c_;
        st.lin = it.lin_base + (int64_t)mlo * g.sM + rc;
        st.pz = a.z_out + it.z_base + (int64_t)it.m0 * g.sM + rc;
    }
    if (XPREV && a.norms_x != nullptr) {
        const int rl = tid / C::TXL, r = it.r0 + rl, c = it.c0 + (tid - rl * C::TXL) * VEC;
        if (r < g.nR && c < g.nC) {
            const PxbVec<T, VEC> v = pxb_vload<T, VEC>(a.x_out + it.lin_base + (int64_t)it.m0 * g.sM + (int64_t)r * g.sR + c);
            for (int j = 0; j < VEC; ++j) st.xp[j] = v.v[j];
        }
    }
    if (q.cp[0] == T(0)) return;  // (run-time test: executed once per work item)
    const int mp = mlo - 1;
    if (!((mp >= 0 || g.open_lo) && (mp < g.nM || g.open_hi))) return;
    const T* __restrict__ z0 = a.z_in + it.z_base + (int64_t)mp * g.sM;
    const int rl = tid / C::TXL, cl = (tid - rl * C::TXL) * VEC;
    {
        const int r = it.r0 + rl, c = it.c0 + cl;
        if (r < g.nR && c < g.nC) { const PxbVec<T, VEC> v = pxb_vload<T, VEC>(z0 + (int64_t)r * g.sR + c); for (int j = 0; j < VEC; ++j) st.z0p[j] = v.v[j]; }
    }
    if (tid < 2 * C::TXL) {
        const int r = pxb_rim_src(tid < C::TXL ? it.r0 - 1 : it.r0 + TY, g.nR, q.mode[1], it.r0, TY, MODES), c = it.c0 + cl;
        if (r != PXB_NOSRC && c < g.nC) { const PxbVec<T, VEC> v = pxb_vload<T, VEC>(z0 + (int64_t)r * g.sR + c); for (int j = 0; j < VEC; ++j) st.z0p_rim[j] = v.v[j]; }
    }
    if (tid >= C::NT - 2 * TY) {
        const int h = tid - (C::NT - 2 * TY);
        const bool left = h < TY;
        const int r = it.r0 + (left ? h : h - TY), c = pxb_rim_src(left ? it.c0 - 1 : it.c0 + C::T2, g.nC, q.mode[2], it.c0, C::T2, MODES);
        if (r < g.nR && c != PXB_NOSRC) st.z0p_col = z0[(int64_t)r * g.sR + c];
    }
}

// PD3O + RelError[x]: the previous x of this thread's samples on the plane phase A visits NEXT (plane m + 1 after phase A of plane m),
// loaded one plane ahead straight into th.xp (a load into a temporary moved over afterwards made the move wait for the load: 9.5 ms
// instead of 7.3 at 1024^3).  Called at the end of phase A's own-sample block.
template <class T, int VEC, int TY, int ALGO, bool NORMS>
PXB_HD void pxb_tma_xprefetch(const PxbIterGeom& g, const PxbIterItem& it, const PxbIterPtr<T>& a, int tid, int m, PxbTmaThread<T, VEC>& th) {
    using C = PxbTmaCfg<T, VEC, TY>;
    if (!(ALGO == PXB_PD3O && NORMS) || PXB_EXP(8) || a.norms_x == nullptr) return;
    if (!(m + 1 >= it.m0 && m + 1 < it.m1)) return;
    const int rl = tid / C::TXL, r = it.r0 + rl, c = it.c0 + (tid - rl * C::TXL) * VEC;
    if (!(it.full || (r < g.nR && c < g.nC))) return;
    const PxbVec<T, VEC> nx = pxb_vload<T, VEC>(a.x_out + th.lin);  // (th.lin: already advanced to plane m + 1)
    for (int j = 0; j < VEC; ++j) th.xp[j] = nx.v[j];
}

// phase A of plane m out of stage `st` (plane m) and, for two-sided / backward schemes, `st_next` (plane m+1).
// MODES (folding boundary modes): the staged boxes are zero-filled outside the domain, i.e. they carry the 'constant'
// extension.  In-domain samples add the fold terms of K^T z (pxb_tv_fold_kz); the cells of the w ring one step outside
// the domain -- rim rows / columns of the edge tiles, the plane past the last one -- receive w at the sample the
// boundary map folds them onto (pxb_tv_w_outside), which K w then reads as the padded array.
template <class T, int VEC, int TY, int ALGO, bool NORMS, class S = PxbSpecAny, bool MODES = false>
PXB_HD void pxb_tma_phaseA(const PxbTvP<T>& q, const PxbIterGeom& g, const PxbTmaGeom& tg, const PxbIterItem& it, const PxbIterPtr<T>& a, int tid,
                           int m, const T* __restrict__ st, const T* __restrict__ st_next, T* ring, PxbTmaThread<T, VEC>& th,
                           const PxbTmaFold& fb = PxbTmaFold{}) {
    using C = PxbTmaCfg<T, VEC, TY>;
    using R = typename C::Ring;
    T* __restrict__ slot = ring + (m & 3) * R::SLOT;
    // MODES: the image's z for the fold terms of K^T z, or null when neither this tile nor this plane holds a fold target
    const T* __restrict__ zfold = (MODES && (fb.any_rc || m == q.fold_hi[0] || m == q.fold_lo[0])) ? a.z_in + it.b * 3 * g.vol : nullptr;
    const bool plane_in = (m >= 0 || g.open_lo) && (m < g.nM || g.open_hi);
    const bool own = m >= it.m0 && m < it.m1;
    const int rl = tid / C::TXL, cl = (tid - rl * C::TXL) * VEC;
    {
        const int r = it.r0 + rl, c = it.c0 + cl;
        const bool in_rc = it.full || (r < g.nR && c < g.nC);
        const bool in = plane_in && in_rc;
        T wv[VEC], z0c[VEC], xo[VEC], un[VEC], uo[VEC];
        pxb_tma_w<T, VEC, TY, VEC, ALGO, S, MODES>(q, tg, st, st_next, rl + 1, cl + VEC, th.z0p, wv, z0c, th.zc[1], th.zc[2], xo, un, uo,
                                                   in ? zfold : nullptr, m, r, c, &fb);
        for (int j = 0; j < VEC; ++j) { th.zc[0][j] = z0c[j]; th.z0p[j] = z0c[j]; }
        // (on the planes of neighbouring chunks / slabs only the tile's own, in-plane-inside cells are read)
        const bool fold = MODES && !in && (own || !plane_in);
        if (fold) pxb_tv_w_outside<T, VEC, 3, ALGO>(q, a.u_in, a.z_in, it.b, m, r, c, wv);
        const bool keep = fold || in;
        PxbVec<T, VEC> o;
        for (int j = 0; j < VEC; ++j) o.v[j] = wv[j];
        if (!keep)  // (never taken on a full tile and an in-domain plane)
            for (int j = 0; j < VEC; ++j) o.v[j] = T(0);
        pxb_vstore<T, VEC>(slot + (rl + 1) * R::RS + cl + VEC, o);
        const int64_t lin = th.lin;  // (= it.lin_base + m * g.sM + r * g.sR + c)
        th.lin = lin + g.sM;
        if (own && in) {
            if (ALGO == PXB_PD3O) {
                if (NORMS && a.norms_x) {
                    // partial sums of one vector in the working precision, widened once (fp32: the per-sample fp64
                    // conversions and FMAs made the criterion-carrying instance issue-bound: 9.3 ms instead of 7.5 at 1024^3)
                    T s0 = T(0), s1 = T(0);  // (th.xp: loaded one plane ahead -- by the prologue for the work item's first plane)
                    for (int j = 0; j < VEC; ++j) {
                        const T dd = xo[j] - th.xp[j];
                        s0 += dd * dd;
                        s1 += th.xp[j] * th.xp[j];
                    }
                    th.acc[0] += (double)s0;
                    th.acc[1] += (double)s1;
                }
                if (a.x_out) { for (int j = 0; j < VEC; ++j) o.v[j] = xo[j]; pxb_vstore<T, VEC>(a.x_out + lin, o); }
            } else if (NORMS && a.norms_x) {
                T s0 = T(0), s1 = T(0);
                for (int j = 0; j < VEC; ++j) {
                    const T dd = un[j] - uo[j];
                    s0 += dd * dd;
                    s1 += uo[j] * uo[j];
                }
                th.acc[0] += (double)s0;
                th.acc[1] += (double)s1;
            }
            for (int j = 0; j < VEC; ++j) o.v[j] = un[j];
            pxb_vstore<T, VEC>(a.u_out + lin, o);
            // peer-memory exchange: the first owned plane of the new primal iterate is the lower neighbour's upper ghost plane
            if (!it.nopeer && m == 0 && a.peer.dn_u != nullptr) pxb_vstore<T, VEC>(a.peer.dn_u + (int64_t)r * g.sR + c, o);
        }
        pxb_tma_xprefetch<T, VEC, TY, ALGO, NORMS>(g, it, a, tid, m, th);
    }
    // rims: w of the neighbouring tiles' border samples (only on planes this work item updates); the carried z0 of
    // the previous plane is refreshed on every plane
    if (tid < 2 * C::TXL) {
        const bool top = tid < C::TXL;
        const int br = top ? 0 : TY + 1;
        const bool need = top ? pxb_has_cm<S>(q, 1) : pxb_has_cp<S>(q, 1);
        int rs = 0, brs = br;  // MODES: the row this rim row stands for (pxb_rim_src) and its row in the staged boxes
        if (MODES) {
            rs = pxb_rim_src(top ? it.r0 - 1 : it.r0 + TY, g.nR, q.mode[1], it.r0, TY, true);
            if (rs != PXB_NOSRC) brs = rs - (it.r0 - 1);
        }
        if (own && need && !PXB_EXP(32)) {
            const int r = top ? it.r0 - 1 : it.r0 + TY, c = it.c0 + cl;
            T wv[VEC], z0c[VEC], z1c[VEC], z2c[VEC], xo[VEC], un[VEC], uo[VEC];
            if (MODES) {
                const bool ev = rs != PXB_NOSRC && c < g.nC;
                if (ev) pxb_tma_w<T, VEC, TY, VEC, ALGO, S, true>(q, tg, st, st_next, brs, cl + VEC, th.z0p_rim, wv, z0c, z1c, z2c, xo, un, uo,
                                                                  zfold, m, rs, c, &fb);
                else pxb_tv_w_outside<T, VEC, 3, ALGO>(q, a.u_in, a.z_in, it.b, m, r, c, wv);  // a fold onto another tile ('wrap'), or zeros
                PxbVec<T, VEC> o;
                for (int j = 0; j < VEC; ++j) o.v[j] = wv[j];
                pxb_vstore<T, VEC>(slot + br * R::RS + cl + VEC, o);
            } else {
                pxb_tma_w<T, VEC, TY, VEC, ALGO, S>(q, tg, st, st_next, br, cl + VEC, th.z0p_rim, wv, z0c, z1c, z2c, xo, un, uo);
                const bool in = r >= 0 && r < g.nR && c < g.nC;
                PxbVec<T, VEC> o;
                for (int j = 0; j < VEC; ++j) o.v[j] = in ? wv[j] : T(0);
                pxb_vstore<T, VEC>(slot + br * R::RS + cl + VEC, o);
            }
        }
        const PxbVec<T, VEC> z = pxb_vload<T, VEC>(st + C::OFF_Z0 + (MODES ? brs : br) * C::BW + cl + VEC);
        for (int j = 0; j < VEC; ++j) th.z0p_rim[j] = pxb_landed(q, z.v[j]);
    }
    if (tid >= C::NT - 2 * TY) {
        const int h = tid - (C::NT - 2 * TY);
        const bool left = h < TY;
        const int hl = left ? h : h - TY;
        const int bc = left ? VEC - 1 : VEC + C::T2;
        const bool need = left ? pxb_has_cm<S>(q, 2) : pxb_has_cp<S>(q, 2);
        int cs = 0, bcs = bc;  // MODES: the column this rim cell stands for and its column in the staged boxes
        if (MODES) {
            cs = pxb_rim_src(left ? it.c0 - 1 : it.c0 + C::T2, g.nC, q.mode[2], it.c0, C::T2, true);
            if (cs != PXB_NOSRC) bcs = cs - (it.c0 - VEC);
        }
        if (own && need) {
            const int r = it.r0 + hl, c = left ? it.c0 - 1 : it.c0 + C::T2;
            T wv[1], z0c[1], z1c[1], z2c[1], xo[1], un[1], uo[1];
            if (MODES) {
                const bool ev = r < g.nR && cs != PXB_NOSRC;
                if (ev) pxb_tma_w<T, VEC, TY, 1, ALGO, S, true>(q, tg, st, st_next, hl + 1, bcs, &th.z0p_col, wv, z0c, z1c, z2c, xo, un, uo,
                                                                zfold, m, r, cs, &fb);
                else pxb_tv_w_outside<T, 1, 3, ALGO>(q, a.u_in, a.z_in, it.b, m, r, c, wv);
                slot[(hl + 1) * R::RS + bc] = wv[0];
            } else {
                pxb_tma_w<T, VEC, TY, 1, ALGO, S>(q, tg, st, st_next, hl + 1, bc, &th.z0p_col, wv, z0c, z1c, z2c, xo, un, uo);
                const bool in = r < g.nR && c >= 0 && c < g.nC;
                slot[(hl + 1) * R::RS + bc] = in ? wv[0] : T(0);
            }
        }
        th.z0p_col = pxb_landed(q, st[C::OFF_Z0 + (hl + 1) * C::BW + (MODES ? bcs : bc)]);
    }
}
