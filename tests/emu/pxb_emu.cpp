// tests/emu/pxb_emu.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the per-voxel bodies of pyxu_b200/csrc/pxb_core.cuh (the exact functions the CUDA
// kernels call) for the host and drives them with plain loops, so that index handling (boundary
// maps, adjoint pre-images, slab halos, fused half-steps) can be checked against the oracle on the
// GPU-less build container.  Never loaded by the pyxu_b200 package; pointers here are HOST pointers.
#include <algorithm>
#include <cstdint>
#include <cmath>
#include <cstring>

static long g_w_global_cells = 0;  // cells of the w tiles evaluated by pxb_tv_w_global (folded rims that left the tile)
#define PXB_EMU_COUNT_W_GLOBAL g_w_global_cells
#include "../../pyxu_b200/csrc/pxb_core.cuh"
#include "../../pyxu_b200/csrc/pxb_tv_fast.cuh"
#include "../../pyxu_b200/csrc/pxb_tv_iter.cuh"
#include "../../pyxu_b200/csrc/pxb_tv_tma.cuh"
#include "../../pyxu_b200/csrc/pxb_stencil_tma.cuh"
#include "../../pyxu_b200/csrc/pxb_tv_tile2d.cuh"
#include "../../pyxu_b200/csrc/pxb_stencil3d.cuh"
#include "../../pyxu_b200/csrc/pxb_stencil3d_dense.cuh"
#include "../../pyxu_b200/csrc/pxb_stencil_axis0.cuh"
#include <vector>

static int g_d3_chunk = 0;  // emu_stencil3d_dense: forced chunk length (0: the launcher's choice)

// folding boundary modes inside the single-kernel iteration forms (pxb_set_iter_modes on the device side)
static bool g_allow_modes = true;

#define FOR_VOX(batch, g)                              \
    for (int64_t b = 0; b < (batch); ++b)              \
        for (int i0 = 0; i0 < (g).n0; ++i0)            \
            for (int i1 = 0; i1 < (g).n1; ++i1)        \
                for (int i2 = 0; i2 < (g).n2; ++i2)

template <class T>
static void t_stencil(const pxb_stencil_desc* d, bool adj, const void* in, void* out) {
    const PxbGeom g = pxb_geom(d->shape);
    FOR_VOX(d->batch, g) pxb_body_stencil<T>(*d, g, adj, (const T*)in, (T*)out, b, i0, i1, i2);
}
template <class T>
static void t_grad(const pxb_grad_desc* d, bool adj, const void* in, void* out) {
    const PxbGeom g = pxb_geom(d->shape);
    FOR_VOX(d->batch, g) {
        if (adj) pxb_body_grad_adjoint<T>(*d, g, (const T*)in, (T*)out, b, i0, i1, i2);
        else pxb_body_grad_apply<T>(*d, g, (const T*)in, (T*)out, b, i0, i1, i2);
    }
}
template <class T>
static void t_primal(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, void* xu, const void* z, const void* ktz, void* x_out,
                     void* w, double* norms) {
    const PxbGeom g = pxb_geom(K->shape);
    // kernels read z / x_out(old) / xu(old) of *other or same* voxels while writing xu, x_out, w: only z has
    // neighbour reads and z is never written here, so a sequential sweep is equivalent to the parallel one.
    FOR_VOX(K->batch, g) {
        double a0 = 0, a1 = 0;
        pxb_body_primal<T>(algo, *K, g, *p, (T*)xu, (const T*)z, (const T*)ktz, (T*)x_out, (T*)w, norms != nullptr, a0, a1, b, i0, i1, i2);
        if (norms) { norms[2 * b] += a0; norms[2 * b + 1] += a1; }
    }
}
template <class T>
static void t_dual(const pxb_grad_desc* K, const pxb_pds_params* p, const void* w, void* z, double* norms) {
    const PxbGeom g = pxb_geom(K->shape);
    FOR_VOX(K->batch, g) {
        double a0 = 0, a1 = 0;
        pxb_body_dual<T>(*K, g, *p, (const T*)w, (T*)z, norms != nullptr, a0, a1, b, i0, i1, i2);
        if (norms) { norms[2 * b] += a0; norms[2 * b + 1] += a1; }
    }
}

// fast (vectorised) TV bodies: loop over vectors of VEC voxels exactly like the CUDA thread map does
template <class T, int NDIR, int VEC>
static void t_tv(int which, int algo, const pxb_grad_desc* K, const PxbTvCoef& cf, const pxb_pds_params* p, void* xu, const void* z,
                 void* x_out, void* w, double* norms) {
    const PxbGeom g = pxb_geom(K->shape);
    PxbTvP<T> q;
    pxb_tv_prepare<T>(*K, cf, *p, q);
    for (int64_t b = 0; b < K->batch; ++b)
        for (int i0 = 0; i0 < g.n0; ++i0)
            for (int i1 = 0; i1 < g.n1; ++i1)
                for (int i2 = 0; i2 < g.n2; i2 += VEC) {
                    double a[2] = {0, 0};
                    if (which == 0) {
                        if (algo == PXB_PD3O) {
                            if (norms) pxb_tv_primal_vec<T, NDIR, VEC, PXB_PD3O, true>(q, *K, *p, (T*)xu, (const T*)z, (T*)x_out, (T*)w, a, b, i0, i1, i2);
                            else pxb_tv_primal_vec<T, NDIR, VEC, PXB_PD3O, false>(q, *K, *p, (T*)xu, (const T*)z, (T*)x_out, (T*)w, a, b, i0, i1, i2);
                        } else {
                            if (norms) pxb_tv_primal_vec<T, NDIR, VEC, PXB_CV, true>(q, *K, *p, (T*)xu, (const T*)z, (T*)x_out, (T*)w, a, b, i0, i1, i2);
                            else pxb_tv_primal_vec<T, NDIR, VEC, PXB_CV, false>(q, *K, *p, (T*)xu, (const T*)z, (T*)x_out, (T*)w, a, b, i0, i1, i2);
                        }
                    } else {
                        if (norms) pxb_tv_dual_vec<T, NDIR, VEC, true>(q, *K, *p, (const T*)w, (T*)z, a, b, i0, i1, i2);
                        else pxb_tv_dual_vec<T, NDIR, VEC, false>(q, *K, *p, (const T*)w, (T*)z, a, b, i0, i1, i2);
                    }
                    if (norms) { norms[2 * b] += a[0]; norms[2 * b + 1] += a[1]; }
                }
}
template <class T, int NDIR>
static int t_tv_vec(int vec, int which, int algo, const pxb_grad_desc* K, const PxbTvCoef& cf, const pxb_pds_params* p, void* xu,
                    const void* z, void* x_out, void* w, double* norms) {
    if (K->shape[2] % vec) return -1;
    if (vec == 4) t_tv<T, NDIR, 4>(which, algo, K, cf, p, xu, z, x_out, w, norms);
    else if (vec == 2) t_tv<T, NDIR, 2>(which, algo, K, cf, p, xu, z, x_out, w, norms);
    else if (vec == 1) t_tv<T, NDIR, 1>(which, algo, K, cf, p, xu, z, x_out, w, norms);
    else return -1;
    return 0;
}
template <class T>
static int t_tv_dir(int vec, int which, int algo, const pxb_grad_desc* K, const pxb_pds_params* p, void* xu, const void* z, void* x_out,
                    void* w, double* norms) {
    PxbTvCoef cf;
    if (!pxb_tv_fast_coefs(*K, cf)) return -2;
    if (K->ndir == 3) return t_tv_vec<T, 3>(vec, which, algo, K, cf, p, xu, z, x_out, w, norms);
    if (K->ndir == 2) return t_tv_vec<T, 2>(vec, which, algo, K, cf, p, xu, z, x_out, w, norms);
    return t_tv_vec<T, 1>(vec, which, algo, K, cf, p, xu, z, x_out, w, norms);
}

// single-kernel iteration (pxb_tv_iter.cuh): every CTA is replayed as "phase A for all threads, then phase C for
// all threads" per plane -- the order the device's one barrier per plane enforces -- with the CTA's shared-memory
// ring as a host array.  Same template instances as the launcher in pxb_tv_iter.cu.
template <class T, int VEC, int TXL, int TY, int NDIR, int ALGO, bool NORMS, bool MODES>
static int t_iter_cfg(const pxb_grad_desc* K, const pxb_pds_params* p, const PxbIterPtr<T>& a, int chunk) {
    using C = PxbIterCfg<T, VEC, TXL, TY, NDIR>;
    PxbTvCoef cf;
    PxbIterGeom g;
    if (int why = pxb_iter_setup(*K, *p, VEC, TY, C::T2, chunk, 5, cf, g, g_allow_modes)) return -100 - why;
    if (pxb_any_mode(*K) != MODES) return -130;  // the launcher's rule: folding modes <=> the MODES instance
    PxbTvP<T> q;
    pxb_tv_prepare<T>(*K, cf, *p, q);
    std::vector<T> smem(C::NSLOT * C::SLOT);
    std::vector<PxbIterThread<T, VEC>> st(C::NT);
    for (int64_t blk = 0; blk < g.nblocks; ++blk) {
        const PxbIterItem it = pxb_iter_item(g, blk, TY, C::T2);
        const PxbIterRange R = pxb_iter_range<T>(q, it);
        for (auto& s : st) std::memset(&s, 0, sizeof(s));
        for (auto& v : smem) v = T(12345);  // poison: cells that are read must have been written
        for (int m = R.mlo; m < R.mhi; ++m) {
            for (int tid = 0; tid < C::NT; ++tid) pxb_iter_phaseA<T, VEC, TXL, TY, NDIR, ALGO, NORMS, MODES>(q, g, it, a, tid, m, smem.data(), st[tid]);
            const int mm = m - R.lag;
            for (int tid = 0; tid < C::NT; ++tid) {
                if (mm >= it.m0 && mm < it.m1)
                    pxb_iter_phaseC<T, VEC, TXL, TY, NDIR, NORMS>(q, g, it, a, tid, mm, smem.data(), R.lag ? st[tid].zprev : st[tid].zc, st[tid].acc);
                std::memcpy(st[tid].zprev, st[tid].zc, sizeof(st[tid].zc));
            }
        }
        if (NORMS)
            for (int tid = 0; tid < C::NT; ++tid) {
                if (a.norms_x) { a.norms_x[2 * it.b] += st[tid].acc[0]; a.norms_x[2 * it.b + 1] += st[tid].acc[1]; }
                if (a.norms_z) { a.norms_z[2 * it.b] += st[tid].acc[2]; a.norms_z[2 * it.b + 1] += st[tid].acc[3]; }
            }
    }
    return 0;
}
template <class T, int NDIR, int ALGO, bool NORMS, bool MODES>
static int t_iter_tile_m(const pxb_grad_desc* K, const pxb_pds_params* p, const PxbIterPtr<T>& a, int chunk) {
    constexpr int VEC = 16 / (int)sizeof(T);
    if (NDIR == 3) return t_iter_cfg<T, VEC, 32, 8, 3, ALGO, NORMS, MODES>(K, p, a, chunk);
    if (K->shape[2] <= 128 * VEC) return t_iter_cfg<T, VEC, 128, 1, 2, ALGO, NORMS, MODES>(K, p, a, chunk);
    return t_iter_cfg<T, VEC, 256, 1, 2, ALGO, NORMS, MODES>(K, p, a, chunk);
}
template <class T, int NDIR, int ALGO, bool NORMS>
static int t_iter_tile(const pxb_grad_desc* K, const pxb_pds_params* p, const PxbIterPtr<T>& a, int chunk) {
    const int rc = t_iter_tile_m<T, NDIR, ALGO, NORMS, false>(K, p, a, chunk);
    return rc != -130 ? rc : t_iter_tile_m<T, NDIR, ALGO, NORMS, true>(K, p, a, chunk);
}
template <class T>
static int t_iter(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, const void* u_in, const void* z_in, void* u_out, void* z_out,
                  void* x_out, double* nx, double* nz, int chunk) {
    PxbIterPtr<T> a{(const T*)u_in, (const T*)z_in, (T*)u_out, (T*)z_out, (T*)x_out, nx, nz};
    const bool norms = nx || nz;
#define EMU_ITER_CASE(ND, AL) \
    if (K->ndir == ND && algo == AL) return norms ? t_iter_tile<T, ND, AL, true>(K, p, a, chunk) : t_iter_tile<T, ND, AL, false>(K, p, a, chunk);
    EMU_ITER_CASE(3, PXB_PD3O)
    EMU_ITER_CASE(3, PXB_CV)
    EMU_ITER_CASE(2, PXB_PD3O)
    EMU_ITER_CASE(2, PXB_CV)
#undef EMU_ITER_CASE
    return -102;
}

// TMA-staged form (pxb_tv_tma.cuh): the box loads the device issues as cp.async.bulk.tensor are replayed as a
// gather with zero fill from the same PxbTmaBoxDesc the tensor-map encoder consumes; the rest is the device's
// per-thread code.  Stage reuse follows the kernel: plane m lives in stage (m - mlo) % 3.
template <class T>
static void emu_tma_box(const PxbTmaBoxDesc& m, uint32_t rows, const int c[5], T* dst) {
    const T* base = (const T*)m.base;
    for (uint32_t i1 = 0; i1 < rows; ++i1)
        for (uint32_t i0 = 0; i0 < m.box[0]; ++i0) {
            const int64_t x0 = (int64_t)c[0] + i0, x1 = (int64_t)c[1] + i1;
            bool in = x0 >= 0 && x0 < (int64_t)m.dim[0] && x1 >= 0 && x1 < (int64_t)m.dim[1];
            for (int k = 2; k < 5; ++k) in = in && c[k] >= 0 && c[k] < (int64_t)m.dim[k];
            T v = T(0);
            if (in) v = base[x0 + x1 * (int64_t)m.stride[1] + c[2] * (int64_t)m.stride[2] + c[3] * (int64_t)m.stride[3] + c[4] * (int64_t)m.stride[4]];
            dst[i1 * m.box[0] + i0] = v;
        }
}
static const pxb_peer* g_emu_peer = nullptr;  // set by emu_tv_iter_tma_p2p for the duration of one call
static int g_emu_edge_first = 1;              // block order of the peer-exchange launches (emu_set_edge_first: 1 = edges first, 2 = interleaved)
template <class T, int ALGO, bool NORMS, class S, bool MODES = false>
static int t_tma_run(const pxb_grad_desc* K, const pxb_pds_params* p, const PxbIterPtr<T>& a_in, int chunk, int want_spec) {
    constexpr int VEC = 16 / (int)sizeof(T), TY = 8;
    using C = PxbTmaCfg<T, VEC, TY>;
    using R = typename C::Ring;
    PxbTvCoef cf;
    PxbIterGeom g;
    PxbIterPtr<T> a = a_in;
    if (int why = pxb_iter_setup(*K, *p, VEC, TY, C::T2, chunk, 5, cf, g, g_allow_modes)) return -100 - why;
    if (pxb_any_mode(*K) != MODES) return -130;  // folding modes <=> the MODES instances, as in the launcher
    PxbTvP<T> q;
    pxb_tv_prepare<T>(*K, cf, *p, q);
    PxbTmaGeom tg;
    PxbTmaBoxDesc mu, ms, mz;
    if (int why = pxb_tma_setup<T, VEC, TY>(*K, *p, cf, g, q, a.u_in, a.z_in, tg, mu, ms, mz)) return -100 - why;
    {   // the caller dispatches on the same rule as the launcher (MODES: specialised for the PD3O-style instances 1 / 2 only)
        int spec = pxb_tma_pick_spec<T>(cf, q, tg);
        if (MODES && spec > 2) spec = 0;
        if (spec != want_spec) return -130;
    }
    if (const pxb_peer* peer = g_emu_peer) {  // as run() in pxb_tv_tma.cu
        if (K->batch != 1) return -124;
        g.edge_first = g_emu_edge_first;
        a.peer.dn_u = (T*)peer->dn_u; a.peer.dn_z = (T*)peer->dn_z; a.peer.dn_zvol = peer->dn_zvol; a.peer.up_z0 = (T*)peer->up_z0;
        a.peer.dn_flag = peer->dn_flag; a.peer.up_flag = peer->up_flag; a.peer.lo_wait = peer->lo_wait; a.peer.hi_wait = peer->hi_wait;
        a.peer.target = (unsigned)((uint64_t)peer->epoch * (uint64_t)g.ntR * (uint64_t)g.ntC);
    }
    std::vector<T> stages(C::NSTAGE * C::STAGE), ring(R::NSLOT * R::SLOT);
    std::vector<PxbTmaThread<T, VEC>> th(C::NT);
    for (int64_t blk = 0; blk < g.nblocks; ++blk) {
        const PxbIterItem it = pxb_iter_item(g, blk, TY, C::T2);
        // the waits of the kernel's prologue: in this sequential replay the neighbour's previous iteration is long over, so a
        // counter short of its target is a counting error
        if (a.peer.lo_wait && it.m0 == 0 && (int)(*a.peer.lo_wait - a.peer.target) < 0) return -140;
        if (a.peer.hi_wait && it.m1 == g.nM && (int)(*a.peer.hi_wait - a.peer.target) < 0) return -141;
        {   // the edge work items come early: all of them first (1), or inside the first 4 x (their number) blocks (2)
            const int64_t eb = (int64_t)(g.nchunk >= 2 ? 2 : 1) * g.ntR * g.ntC;
            const bool is_edge = it.m0 == 0 || it.m1 == g.nM;
            const bool woven = g.edge_first == 2 && (int64_t)g.ntR * g.ntC * g.nchunk - eb >= 3 * eb;
            if (g.edge_first && !woven && blk < eb && !is_edge) return -142;
            if (g.edge_first && woven && blk >= 4 * eb && is_edge) return -142;
        }
        const PxbIterRange Rg = pxb_iter_range<T>(q, it);
        const bool need_next = pxb_has_cm<S>(q, 0);
        const int lag = S::SCHEME == PXB_SCHEME_FWD ? 1 : Rg.lag;
        const int mload_hi = Rg.mhi + (need_next ? 1 : 0);
        for (auto& v : stages) v = T(777);
        for (auto& v : ring) v = T(12345);
        auto issue = [&](int m) {
            T* st = stages.data() + ((m - Rg.mlo) % C::NSTAGE) * C::STAGE;
            const int b = (int)it.b;
            int cu[5] = {it.c0 - VEC, it.r0 - 1, m + tg.gl, 0, b};
            emu_tma_box<T>(mu, C::BR, cu, st + C::OFF_U);
            if (S::FK >= 1 || tg.has_shift) { int cs[5] = {cu[0], cu[1], cu[2], 0, tg.sh_batched ? b : 0}; emu_tma_box<T>(ms, C::BR, cs, st + C::OFF_S); }
            emu_tma_box<T>(mz, C::BR, cu, st + C::OFF_Z0);
            int c2[5] = {cu[0], cu[1], cu[2], 2, b};
            emu_tma_box<T>(mz, C::BR, c2, st + C::OFF_Z2);
            int c1[5] = {cu[0], cu[1] - 1, cu[2], 1, b};
            emu_tma_box<T>(mz, C::BR1, c1, st + C::OFF_Z1);
        };
        for (int m = Rg.mlo; m < Rg.mlo + C::NSTAGE && m < mload_hi; ++m) issue(m);
        const PxbTmaFold fbx = MODES ? pxb_tma_fold_setup<T, VEC, TY>(q, g, it) : PxbTmaFold{};
        for (int tid = 0; tid < C::NT; ++tid) {
            std::memset(&th[tid], 0, sizeof(th[tid]));
            pxb_tma_prologue<T, VEC, TY, MODES, ALGO == PXB_PD3O && NORMS>(q, g, it, a, tid, Rg.mlo, th[tid]);
        }
        for (int m = Rg.mlo; m < Rg.mhi; ++m) {
            const int k = m - Rg.mlo;
            const T* st = stages.data() + (k % C::NSTAGE) * C::STAGE;
            const T* st_next = need_next ? stages.data() + ((k + 1) % C::NSTAGE) * C::STAGE : st;
            for (int tid = 0; tid < C::NT; ++tid) pxb_tma_phaseA<T, VEC, TY, ALGO, NORMS, S, MODES>(q, g, tg, it, a, tid, m, st, st_next, ring.data(), th[tid], fbx);
            if (m + C::NSTAGE < mload_hi) issue(m + C::NSTAGE);
            const int mm = m - lag;
            for (int tid = 0; tid < C::NT; ++tid) {
                if (mm >= it.m0 && mm < it.m1) {
                    pxb_iter_phaseC<T, VEC, C::TXL, TY, 3, NORMS, S>(q, g, it, a, tid, mm, ring.data(), lag ? th[tid].zprev : th[tid].zc, th[tid].acc, th[tid].pz);
                    th[tid].pz += g.sM;
                }
                std::memcpy(th[tid].zprev, th[tid].zc, sizeof(th[tid].zc));
            }
        }
        if (a.peer.dn_flag && it.m0 == 0) *a.peer.dn_flag += 1u;
        if (a.peer.up_flag && it.m1 == g.nM) *a.peer.up_flag += 1u;
        if (NORMS)
            for (int tid = 0; tid < C::NT; ++tid) {
                if (a.norms_x) { a.norms_x[2 * it.b] += th[tid].acc[0]; a.norms_x[2 * it.b + 1] += th[tid].acc[1]; }
                if (a.norms_z) { a.norms_z[2 * it.b] += th[tid].acc[2]; a.norms_z[2 * it.b + 1] += th[tid].acc[3]; }
            }
    }
    return 0;
}
template <class T>
static int t_tma(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, const void* u_in, const void* z_in, void* u_out, void* z_out,
                 void* x_out, double* nx, double* nz, int chunk) {
    if (K->ndir != 3) return -120;
    PxbIterPtr<T> a{(const T*)u_in, (const T*)z_in, (T*)u_out, (T*)z_out, (T*)x_out, nx, nz};
    const bool norms = nx || nz;
    // try the instances in the launcher's order; exactly one accepts (-130 = "not my problem")
#define EMU_TMA_TRY_M(S, id, M)                                                                                         \
    {                                                                                                                   \
        int rc = algo == PXB_PD3O ? (norms ? t_tma_run<T, PXB_PD3O, true, S, M>(K, p, a, chunk, id) : t_tma_run<T, PXB_PD3O, false, S, M>(K, p, a, chunk, id)) \
                                  : (norms ? t_tma_run<T, PXB_CV, true, S, M>(K, p, a, chunk, id) : t_tma_run<T, PXB_CV, false, S, M>(K, p, a, chunk, id));       \
        if (rc != -130) return rc;                                                                                      \
    }
#define EMU_TMA_TRY(S, id) EMU_TMA_TRY_M(S, id, false)
    EMU_TMA_TRY_M(PxbSpecFwdPos, 1, true)
    EMU_TMA_TRY_M(PxbSpecFwdNone, 2, true)
    EMU_TMA_TRY_M(PxbSpecAny, 0, true)
    EMU_TMA_TRY(PxbSpecFwdPos, 1)
    EMU_TMA_TRY(PxbSpecFwdNone, 2)
    if (algo == PXB_CV) {
        EMU_TMA_TRY(PxbSpecFwdPosG, 3)
        EMU_TMA_TRY(PxbSpecFwdNoneG, 4)
    }
    EMU_TMA_TRY(PxbSpecAny, 0)
#undef EMU_TMA_TRY
#undef EMU_TMA_TRY_M
    return -131;
}

// TMA-tiled 2-D stencil (pxb_stencil_tma.cuh): box load emulated as a zero-filled gather, then the device's
// per-thread bodies in barrier order.
template <class T, int VEC, int NV>
static void t_st2_nv(const PxbSt2P& p, const PxbSt2In& ext, const T* in, const T* in2, T* out) {
    const int in_n1 = ext.n1, in_n2 = ext.n2;
    using C = PxbSt2Cfg<T, VEC>;
    std::vector<T> box((size_t)p.bh * p.bw), box2((size_t)p.bh * p.bw), mid((size_t)p.bh * C::TX + (size_t)p.k1 * p.k2);
    for (int64_t img = 0; img < p.nimg; ++img)
        for (int ty = 0; ty < p.nty; ++ty)
            for (int tx = 0; tx < p.ntx; ++tx) {
                const int x0 = tx * C::TX, y0 = ty * C::TY;
                for (int i = 0; i < p.bh; ++i)
                    for (int j = 0; j < p.bw; ++j) {
                        const int y = y0 - p.c1 + i, x = x0 - p.c2 + j;
                        const bool ok = y >= 0 && y < in_n1 && x >= 0 && x < in_n2;  // the tensor map's extent: zero fill beyond
                        box[(size_t)i * p.bw + j] = ok ? in[(img * in_n1 + y) * (int64_t)in_n2 + x] : T(0);
                        if (in2) box2[(size_t)i * p.bw + j] = ok ? in2[(img * in_n1 + y) * (int64_t)in_n2 + x] : T(0);
                    }
                if (in2)
                    for (int it = 0; it < p.bh * p.bw / VEC; ++it) pxb_st2_combine_item<T, VEC>(p, box.data(), box2.data(), it);
                if (p.dense)
                    for (int i = 0; i < p.k1 * p.k2; ++i) mid[i] = pxb_st2_dense_coef<T>(p, (const T*)p.coef, i);
                T c1[PXB_ST2_MAXTAP + 2 * (C::R - 1)], c2[PXB_ST2_MAXTAP];
                for (int q = 0; q < PXB_ST2_MAXTAP; ++q) c2[q] = T(p.coef2[q]);
                for (int t = 0; t < PXB_ST2_MAXTAP + 2 * (C::R - 1); ++t) {
                    const int q = t - (C::R - 1);
                    c1[t] = (q >= 0 && q < p.k1) ? T(p.coef1[q]) : T(0);
                }
                if (!p.dense)
                    for (int it = 0; it < p.bh * C::TXL; ++it) pxb_st2_row_item<T, VEC, NV>(p, box.data(), mid.data(), it / C::TXL, (it % C::TXL) * VEC, c2);
                for (int tid = 0; tid < C::NT; ++tid) {
                    const int xl = (tid & 31) * VEC, yl = (tid >> 5) * C::R;
                    T acc[C::R][VEC];
                    if (p.dense) pxb_st2_dense_item<T, VEC, NV>(p, box.data(), mid.data(), yl, xl, acc);
                    else pxb_st2_col_item<T, VEC>(p, mid.data(), yl, xl, c1, acc);
                    PxbSt2Epi<T, VEC> epi;
                    pxb_st2_load_epi<T, VEC>(p, epi, img, y0, x0, yl, xl);
                    if (p.epi == 1) {
                        double nrm[2] = {0.0, 0.0};
                        pxb_st2_store_prox<T, VEC>(p, out, epi, img, y0, x0, yl, xl, acc, nrm);
                        if (p.norms) { p.norms[2 * (img / p.imgs_per_row)] += nrm[0]; p.norms[2 * (img / p.imgs_per_row) + 1] += nrm[1]; }
                    } else {
                        pxb_st2_store<T, VEC>(p, out, epi, img, y0, x0, yl, xl, acc);
                    }
                }
            }
}
template <class T>
static int t_st2(const pxb_stencil2d* d, const pxb_fista_step* f, int which, const void* in0, void* out) {
    constexpr int VEC = 16 / (int)sizeof(T);
    PxbSt2P p;
    p.n1 = (int)d->shape[0]; p.n2 = (int)d->shape[1]; p.nimg = d->nimg;
    p.k1 = d->ksize[0]; p.k2 = d->ksize[1]; p.c1 = d->center[0]; p.c2 = d->center[1];
    if (p.c1 < 0 || p.c1 >= p.k1 || p.c2 < 0 || p.c2 >= p.k2) return -103;
    PxbSt2In ext{p.n1, p.n2};
    const bool own_extent = d->in_shape[0] > 0 && d->in_shape[1] > 0;
    if (own_extent) {  // as the launcher: the origin of the output grid folded into the centers
        if (f) return -107;
        ext.n1 = (int)d->in_shape[0]; ext.n2 = (int)d->in_shape[1];
        p.c1 -= d->origin[0]; p.c2 -= d->origin[1];
    }
    p.dense = d->dense;
    for (int i = 0; i < PXB_ST2_MAXTAP; ++i) { p.coef1[i] = d->coef1[i]; p.coef2[i] = d->coef2[i]; }
    p.coef = d->coef; p.alpha = d->alpha; p.beta = d->beta; p.add = d->add; p.add_period = d->add_period;
    if (d->add && d->add_period > 0 && d->add_period >= d->nimg * d->shape[0] * d->shape[1]) p.add_period = 0;
    p.pa = 1.0; p.pb = 0.0; p.epi = 0; p.e1 = p.e2 = nullptr; p.ea = p.eb = 0.0; p.gkind = 0; p.gp0 = p.gp1 = p.tau = 0.0;
    p.norms = nullptr; p.imgs_per_row = 1;
    const void *in = in0, *in2 = nullptr;
    if (f) {  // same mapping as pxb_stencil2d_fista_try
        if (which == 0) {
            in = f->x;
            if (f->a != 0.0) { in2 = f->x_prev; p.pa = 1.0 + f->a; p.pb = -f->a; }
        } else {
            in = f->r;
            p.epi = 1;
            p.e1 = f->x; p.e2 = f->a != 0.0 ? f->x_prev : nullptr;
            p.ea = 1.0 + f->a; p.eb = -f->a;
            p.gkind = f->g.kind; p.gp0 = f->g.p0; p.gp1 = f->g.p1; p.tau = f->tau;
            p.norms = f->norms; p.imgs_per_row = f->imgs_per_row > 0 ? f->imgs_per_row : 1;
        }
    }
    if (int why = pxb_st2_setup<T, VEC>(p, own_extent ? &ext : nullptr)) return -100 - why;
    switch (pxb_st2_nv(p.k2, VEC)) {
        case 1: t_st2_nv<T, VEC, 1>(p, ext, (const T*)in, (const T*)in2, (T*)out); break;
        case 2: t_st2_nv<T, VEC, 2>(p, ext, (const T*)in, (const T*)in2, (T*)out); break;
        case 3: t_st2_nv<T, VEC, 3>(p, ext, (const T*)in, (const T*)in2, (T*)out); break;
        case 4: t_st2_nv<T, VEC, 4>(p, ext, (const T*)in, (const T*)in2, (T*)out); break;
        case 5: t_st2_nv<T, VEC, 5>(p, ext, (const T*)in, (const T*)in2, (T*)out); break;
        case 6: t_st2_nv<T, VEC, 6>(p, ext, (const T*)in, (const T*)in2, (T*)out); break;
        default: return -101;
    }
    return 0;
}

// TMA-tiled 2-D iteration (pxb_tv_tile2d.cuh): boxes gathered with zero fill, then phase A / phase C per thread.
template <class T>
static void emu_box3(const T* base, const int64_t dim[3], int64_t s1, int64_t s2, int bw, int rows, int c0, int c1, int64_t c2, T* dst) {
    for (int i = 0; i < rows; ++i)
        for (int j = 0; j < bw; ++j) {
            const int64_t x = (int64_t)c0 + j, y = (int64_t)c1 + i;
            const bool in = x >= 0 && x < dim[0] && y >= 0 && y < dim[1] && c2 >= 0 && c2 < dim[2];
            dst[i * bw + j] = in ? base[x + y * s1 + c2 * s2] : T(0);
        }
}
template <class T, int ALGO, bool NORMS, class S, bool MODES = false>
static int t_t2_run(const pxb_grad_desc* K, const pxb_pds_params* p, const PxbIterPtr<T>& a, bool want_fwd) {
    constexpr int VEC = 16 / (int)sizeof(T);
    using C = PxbT2Cfg<T, VEC>;
    PxbTvCoef cf;
    PxbTvP<T> q;
    PxbT2Geom g;
    if (int why = pxb_t2_setup<T, VEC>(*K, *p, cf, q, g, g_allow_modes)) return -100 - why;
    if (pxb_any_mode(*K) != MODES) return -130;
    bool fwd = true;
    for (int k = 0; k < 2; ++k) fwd = fwd && cf.cm[k] == 0.0 && cf.cp[k] != 0.0;
    if ((fwd && q.hkind == PXB_DUAL_L21) != want_fwd) return -130;
    std::vector<T> sm(C::TOTAL);
    const int64_t du[3] = {g.n2, g.n1, g.nimg}, dz[3] = {g.n2, g.n1, g.nimg * 2}, ds[3] = {g.n2, g.n1, g.sh_mode ? g.n0 : g.nimg};
    const T* sptr = q.fkind == PXB_F_GRADARR ? q.garr : q.shift;
    for (int64_t blk = 0; blk < g.nblocks; ++blk) {
        const PxbT2Item it = pxb_t2_item(g, blk, C::TY, C::T2);
        for (auto& v : sm) v = T(4321);
        const int cc = it.c0 - VEC, cr = it.r0 - 1;
        emu_box3<T>(a.u_in, du, g.n2, g.s0, C::BW, C::BR, cc, cr, it.img, sm.data() + C::OFF_U);
        if (g.has_shift) emu_box3<T>(sptr, ds, g.n2, g.s0, C::BW, C::BR, cc, cr, g.sh_mode ? it.i0 : it.img, sm.data() + C::OFF_S);
        const int64_t pz = it.b * 2 * g.n0 + it.i0;
        emu_box3<T>(a.z_in, dz, g.n2, g.s0, C::BW, C::BRZ, cc, cr - 1, pz, sm.data() + C::OFF_ZR);
        emu_box3<T>(a.z_in, dz, g.n2, g.s0, C::BW, C::BR, cc, cr, pz + g.n0, sm.data() + C::OFF_ZC);
        double acc[4] = {0, 0, 0, 0};
        for (int tid = 0; tid < C::NT; ++tid) pxb_t2_phaseA<T, VEC, ALGO, NORMS, S, MODES>(q, g, it, a, tid, sm.data(), acc);
        for (int tid = 0; tid < C::NT; ++tid) pxb_t2_phaseC<T, VEC, NORMS, S>(q, g, it, a, tid, sm.data(), acc);
        if (NORMS) {
            if (a.norms_x) { a.norms_x[2 * it.b] += acc[0]; a.norms_x[2 * it.b + 1] += acc[1]; }
            if (a.norms_z) { a.norms_z[2 * it.b] += acc[2]; a.norms_z[2 * it.b + 1] += acc[3]; }
        }
    }
    return 0;
}
template <class T>
static int t_t2(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, const void* u_in, const void* z_in, void* u_out, void* z_out,
                void* x_out, double* nx, double* nz) {
    if (K->ndir != 2) return -102;
    PxbIterPtr<T> a{(const T*)u_in, (const T*)z_in, (T*)u_out, (T*)z_out, (T*)x_out, nx, nz};
    const bool norms = nx || nz;
    using SF = PxbSpec<PXB_SCHEME_FWD, -1, PXB_DUAL_L21, -1>;
#define EMU_T2_TRY_M(S, fw, M)                                                                                        \
    {                                                                                                                 \
        int rc = algo == PXB_PD3O ? (norms ? t_t2_run<T, PXB_PD3O, true, S, M>(K, p, a, fw) : t_t2_run<T, PXB_PD3O, false, S, M>(K, p, a, fw)) \
                                  : (norms ? t_t2_run<T, PXB_CV, true, S, M>(K, p, a, fw) : t_t2_run<T, PXB_CV, false, S, M>(K, p, a, fw));       \
        if (rc != -130) return rc;                                                                                    \
    }
#define EMU_T2_TRY(S, fw) EMU_T2_TRY_M(S, fw, false)
    EMU_T2_TRY_M(SF, true, true)
    EMU_T2_TRY_M(PxbSpecAny, false, true)
    EMU_T2_TRY(SF, true)
    EMU_T2_TRY(PxbSpecAny, false)
#undef EMU_T2_TRY
#undef EMU_T2_TRY_M
    return -131;
}

// vectorised Gradient apply / adjoint bodies (pxb_tv_fast.cuh), looped like the CUDA thread map
template <class T, int NDIR, int VEC>
static void t_tvgrad(bool adj, const pxb_grad_desc* K, const PxbTvCoef& cf, const void* in, void* out) {
    const PxbGeom g = pxb_geom(K->shape);
    pxb_pds_params P{};
    PxbTvP<T> q;
    pxb_tv_prepare<T>(*K, cf, P, q);
    for (int64_t b = 0; b < K->batch; ++b)
        for (int i0 = 0; i0 < g.n0; ++i0)
            for (int i1 = 0; i1 < g.n1; ++i1)
                for (int i2 = 0; i2 < g.n2; i2 += VEC) {
                    if (adj) pxb_tv_grad_adjoint_vec<T, NDIR, VEC>(q, *K, (const T*)in, (T*)out, b, i0, i1, i2);
                    else pxb_tv_grad_apply_vec<T, NDIR, VEC>(q, *K, (const T*)in, (T*)out, b, i0, i1, i2);
                }
}
template <class T, int NDIR>
static int t_tvgrad_vec(int vec, bool adj, const pxb_grad_desc* K, const PxbTvCoef& cf, const void* in, void* out) {
    if (K->shape[2] % vec) return -1;
    if (vec == 4) t_tvgrad<T, NDIR, 4>(adj, K, cf, in, out);
    else if (vec == 2) t_tvgrad<T, NDIR, 2>(adj, K, cf, in, out);
    else if (vec == 1) t_tvgrad<T, NDIR, 1>(adj, K, cf, in, out);
    else return -1;
    return 0;
}
template <class T>
static int t_tvgrad_dir(int vec, bool adj, const pxb_grad_desc* K, const void* in, void* out) {
    PxbTvCoef cf;
    if (!pxb_tv_fast_coefs(*K, cf)) return -2;
    if (K->ndir == 3) return t_tvgrad_vec<T, 3>(vec, adj, K, cf, in, out);
    if (K->ndir == 2) return t_tvgrad_vec<T, 2>(vec, adj, K, cf, in, out);
    return t_tvgrad_vec<T, 1>(vec, adj, K, cf, in, out);
}

// single-pass separable 3-D stencil (pxb_stencil3d.cuh): CTA by CTA, plane by plane; the per-thread register ring is an array
template <class T, int VEC, int NV, int K0>
static void t_st3_run(const PxbSt3P& p, const T* in, T* out) {
    using C = PxbSt3Cfg<T, VEC>;
    const int64_t s0 = (int64_t)p.s.n1 * p.s.n2;
    std::vector<T> box((size_t)p.s.bh * p.s.bw), mid((size_t)p.s.bh * C::TX);
    struct Ring { T v[K0][C::R][VEC]; };
    std::vector<Ring> rings(C::NT);
    T c1[PXB_ST2_MAXTAP + 2 * (C::R - 1)], c2[PXB_ST2_MAXTAP], c0v[K0];
    for (int q = 0; q < PXB_ST2_MAXTAP; ++q) c2[q] = T(p.s.coef2[q]);
    for (int t = 0; t < PXB_ST2_MAXTAP + 2 * (C::R - 1); ++t) { const int q = t - (C::R - 1); c1[t] = (q >= 0 && q < p.s.k1) ? T(p.s.coef1[q]) : T(0); }
    for (int k = 0; k < K0; ++k) c0v[k] = T(p.coef0[k]);
    for (int64_t b = 0; b < p.batch; ++b)
        for (int ch = 0; ch < p.nchunk; ++ch)
            for (int ty = 0; ty < p.s.nty; ++ty)
                for (int tx = 0; tx < p.s.ntx; ++tx) {
                    const int x0 = tx * C::TX, y0 = ty * C::TY;
                    const int m0 = ch * p.chunk, m1 = std::min(p.n0, m0 + p.chunk);
                    const int pl_lo = m0 - p.c0, pl_hi = m1 + K0 - 1 - p.c0;
                    for (auto& r : rings) std::memset(&r, 0, sizeof(r));
                    for (int pl = pl_lo; pl < pl_hi; ++pl) {
                        const bool have = pl >= -p.lo_planes && pl < p.n0 + p.hi_planes;
                        if (have) {
                            for (int i = 0; i < p.s.bh; ++i)
                                for (int j = 0; j < p.s.bw; ++j) {
                                    const int y = y0 - p.s.c1 + i, x = x0 - p.s.c2 + j;
                                    box[(size_t)i * p.s.bw + j] = (y >= 0 && y < p.s.n1 && x >= 0 && x < p.s.n2) ? in[b * p.vol + (int64_t)pl * s0 + (int64_t)y * p.s.n2 + x] : T(0);
                                }
                            for (int it = 0; it < p.s.bh * C::TXL; ++it) pxb_st3_row_item<T, VEC, NV>(p, box.data(), mid.data(), it / C::TXL, (it % C::TXL) * VEC, c2);
                        }
                        const int q = pl - (K0 - 1 - p.c0);
                        const int u = (pl - pl_lo) % K0;  // ring slot of this plane, as in the unrolled device loop
                        for (int tid = 0; tid < C::NT; ++tid) {
                            const int xl = (tid & 31) * VEC, yl = (tid >> 5) * C::R;
                            T addv[C::R][VEC];
                            if (q >= m0) pxb_st3_load_add<T, VEC>(p, addv, b, q, y0, x0, yl, xl);
                            if (have) pxb_st3_col_item<T, VEC>(p, mid.data(), yl, xl, c1, rings[tid].v[u]);
                            else std::memset(rings[tid].v[u], 0, sizeof(rings[tid].v[u]));
                            if (q >= m0) pxb_st3_store<T, VEC, K0>(p, out, c0v, rings[tid].v, u, addv, b, q, y0, x0, yl, xl);
                        }
                    }
                }
}
template <class T, int VEC, int NV>
static int t_st3_k0(const PxbSt3P& p, const T* in, T* out) {
    switch (p.k0) {
        case 3: t_st3_run<T, VEC, NV, 3>(p, in, out); return 0;
        case 5: t_st3_run<T, VEC, NV, 5>(p, in, out); return 0;
        case 7: t_st3_run<T, VEC, NV, 7>(p, in, out); return 0;
        case 9: t_st3_run<T, VEC, NV, 9>(p, in, out); return 0;
        default: return -111;
    }
}
template <class T>
static int t_st3(const pxb_stencil3d* d, const void* in, void* out) {
    constexpr int VEC = 16 / (int)sizeof(T);
    PxbSt3P p;
    p.s.n1 = (int)d->shape[1]; p.s.n2 = (int)d->shape[2];
    p.s.k1 = d->ksize[1]; p.s.k2 = d->ksize[2]; p.s.c1 = d->center[1]; p.s.c2 = d->center[2];
    for (int i = 0; i < PXB_ST2_MAXTAP; ++i) { p.s.coef1[i] = d->coef1[i]; p.s.coef2[i] = d->coef2[i]; p.coef0[i] = d->coef0[i]; }
    p.s.coef = nullptr; p.s.alpha = d->alpha; p.s.beta = d->beta; p.s.add = d->add; p.s.add_period = d->add_period;
    if (d->add && d->add_period > 0 && d->add_period >= d->batch * d->shape[0] * d->shape[1] * d->shape[2]) p.s.add_period = 0;
    p.s.pa = 1.0; p.s.pb = 0.0; p.s.epi = 0; p.s.e1 = p.s.e2 = nullptr; p.s.norms = nullptr; p.s.imgs_per_row = 1;
    p.n0 = (int)d->shape[0]; p.batch = d->batch;
    const int halo = d->slab.halo;
    const int alloc = d->slab.plane_alloc > 0 ? d->slab.plane_alloc : p.n0 + 2 * halo;
    p.vol = (int64_t)alloc * d->shape[1] * d->shape[2];
    p.k0 = d->ksize[0]; p.c0 = d->center[0];
    p.lo_planes = d->slab.open_lo ? p.c0 : 0;
    p.hi_planes = d->slab.open_hi ? p.k0 - 1 - p.c0 : 0;
    if (int why = pxb_st3_setup<T, VEC>(p)) return -100 - why;
    switch (pxb_st2_nv(p.s.k2, VEC)) {
        case 1: return t_st3_k0<T, VEC, 1>(p, (const T*)in, (T*)out);
        case 2: return t_st3_k0<T, VEC, 2>(p, (const T*)in, (T*)out);
        case 3: return t_st3_k0<T, VEC, 3>(p, (const T*)in, (T*)out);
        case 4: return t_st3_k0<T, VEC, 4>(p, (const T*)in, (T*)out);
        case 5: return t_st3_k0<T, VEC, 5>(p, (const T*)in, (T*)out);
        case 6: return t_st3_k0<T, VEC, 6>(p, (const T*)in, (T*)out);
        default: return -101;
    }
}

// dense K x K x K marching stencil (pxb_stencil3d_dense.cuh): CTA by CTA, plane by plane, the phases between two barriers of
// k_stencil3d_dense replayed for all threads in turn; per-thread registers (accumulator slots, fetched window) are arrays
template <class T, int K>
static int t_d3_run(const pxb_stencil3d_dense* d, const T* in, T* out) {
    using C = PxbD3Cfg<T, K>;
    constexpr int VEC = C::VEC;
    PxbD3P<T, K> p;
    p.n0 = (int)d->shape[0]; p.n1 = (int)d->shape[1]; p.n2 = (int)d->shape[2];
    p.batch = d->batch;
    const int halo = d->slab.halo;
    const int alloc = d->slab.plane_alloc > 0 ? d->slab.plane_alloc : p.n0 + 2 * halo;
    p.vol = (int64_t)alloc * d->shape[1] * d->shape[2];
    p.c0 = d->center[0]; p.c1 = d->center[1]; p.c2 = d->center[2];
    p.lo_planes = d->slab.open_lo ? d->center[0] : 0;
    p.hi_planes = d->slab.open_hi ? d->ksize[0] - 1 - d->center[0] : 0;
    p.alpha = T(d->alpha); p.beta = T(d->beta);
    p.add = (const T*)d->add; p.add_period = d->add_period;
    if (d->add && d->add_period > 0 && d->add_period >= d->batch * d->shape[0] * d->shape[1] * d->shape[2]) p.add_period = 0;
    for (int i = 0; i < K * C::CROW; ++i) p.coef[i] = T(0);
    for (int a = 0; a < K; ++a)
        for (int bb = 0; bb < K; ++bb)
            for (int c = 0; c < K; ++c) {
                const bool in_k = a < d->ksize[0] && bb < d->ksize[1] && c < d->ksize[2];
                p.coef[bb * C::CROW + a * K + c] = in_k ? T(d->coef[((int64_t)a * d->ksize[1] + bb) * d->ksize[2] + c]) : T(0);
            }
    if (int why = pxb_d3_setup<T, K>(p)) return -100 - why;
    if (g_d3_chunk > 0) { p.chunk = std::min(g_d3_chunk, p.n0); p.nchunk = (p.n0 + p.chunk - 1) / p.chunk; }  // tests: several chunks on small volumes
    const int64_t s0 = (int64_t)p.n1 * p.n2;
    struct Regs { T acc[K][C::R][VEC]; T pre[C::NROW * C::NCOL]; };
    std::vector<Regs> regs(C::NT);
    std::vector<T> box0(C::BOX), box1(C::BOX);
    T* box[2] = {box0.data(), box1.data()};
    for (int64_t b = 0; b < p.batch; ++b)
        for (int ch = 0; ch < p.nchunk; ++ch)
            for (int ty = 0; ty < p.nty; ++ty)
                for (int tx = 0; tx < p.ntx; ++tx) {
                    const int x0 = tx * C::TX, y0 = ty * C::TY;
                    const int m0 = ch * p.chunk, m1 = std::min(p.n0, m0 + p.chunk);
                    const int pl_lo = m0 - p.c0, pl_hi = m1 + K - 1 - p.c0;
                    const int ra = std::max(pl_lo, -p.lo_planes), rb = std::min(pl_hi, p.n0 + p.hi_planes);
                    const T* vol = in + b * p.vol;
                    for (auto& r : regs) std::memset(&r, 0, sizeof(r));
                    std::fill(box0.begin(), box0.end(), T(NAN));  // a read of a cell no thread staged poisons the result
                    std::fill(box1.begin(), box1.end(), T(NAN));
                    if (ra < rb)
                        for (int tid = 0; tid < C::NT; ++tid) {
                            pxb_d3_fetch<T, K>(p, vol + (int64_t)ra * s0, y0, x0, tid, regs[tid].pre);
                            pxb_d3_stash<T, K>(regs[tid].pre, box[0], tid);
                        }
                    for (int pl = ra; pl < pl_hi; ++pl) {
                        const bool have = pl < rb, more = pl + 1 < rb;
                        const int k = pl - ra, q = pl - (K - 1 - p.c0);
                        for (int tid = 0; tid < C::NT; ++tid) {
                            Regs& r = regs[tid];
                            const int xl = (tid & 31) * VEC, yl = (tid >> 5) * C::R;
                            if (more) pxb_d3_fetch<T, K>(p, vol + (int64_t)(pl + 1) * s0, y0, x0, tid, r.pre);
                            T addv[C::R][VEC];
                            if (q >= m0) pxb_d3_load_add<T, K>(p, addv, b, q, y0, x0, yl, xl);
                            if (have) {
                                const int a_lo = pl + p.c0 - m1 + 1, a_hi = pl + p.c0 - m0;
                                if (a_lo <= 0 && a_hi >= K - 1) pxb_d3_accum<T, K>(p.coef, box[k & 1], yl, xl, r.acc);
                                else pxb_d3_accum_some<T, K>(p.coef, box[k & 1], yl, xl, r.acc, a_lo, a_hi);
                            }
                            if (q >= m0) pxb_d3_emit<T, K>(p, out, r.acc[K - 1], addv, b, q, y0, x0, yl, xl);
                            pxb_d3_shift<T, K>(r.acc);
                        }
                        if (more)
                            for (int tid = 0; tid < C::NT; ++tid) pxb_d3_stash<T, K>(regs[tid].pre, box[(k + 1) & 1], tid);
                    }
                }
    return 0;
}
template <class T>
static int t_d3(const pxb_stencil3d_dense* d, const void* in, void* out) {
    switch (pxb_d3_cube(d->ksize)) {
        case 3: return t_d3_run<T, 3>(d, (const T*)in, (T*)out);
        case 5: return t_d3_run<T, 5>(d, (const T*)in, (T*)out);
        case 7: return t_d3_run<T, 7>(d, (const T*)in, (T*)out);
        default: return -101;
    }
}

template <class T>
static void t_pad2d(const pxb_pad2d_desc* d, bool adj, const void* a, void* b, double alpha, double beta, const void* add, int64_t add_period) {
    const int64_t rows = adj ? d->shape[0] : d->ext_shape[0], cols = adj ? d->shape[1] : d->ext_shape[1];
    if (add && add_period >= d->nimg * d->shape[0] * d->shape[1]) add_period = 0;
    for (int64_t img = 0; img < d->nimg; ++img)
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) {
                const int64_t lin = (img * rows + r) * cols + c;
                if (!adj) ((T*)b)[lin] = pxb_pad2d_at<T>(*d, (const T*)a, img, r, c);
                else {
                    T o = T(alpha) * pxb_pad2d_adj_at<T>(*d, (const T*)a, img, r, c);
                    if (add) o += T(beta) * ((const T*)add)[add_period > 0 ? lin % add_period : lin];
                    ((T*)b)[lin] = o;
                }
            }
}
// streaming axis-0 stencil with a folding mode (pxb_stencil_axis0.cuh): column by column, chunk by chunk, as the grid does
template <class T, int K0>
static void t_axis0_k(const Axis0P& p, int64_t batch, const T* in, T* out) {
    constexpr int VEC = 16 / (int)sizeof(T);
    for (int64_t b = 0; b < batch; ++b)
        for (int ch = 0; ch < p.nchunk; ++ch)
            for (int64_t col = 0; col < p.plane; col += VEC) {
                const int m0 = ch * p.chunk, m1 = std::min(p.n0, m0 + p.chunk);
                pxb_axis0_column<T, VEC, K0, true>(p, in + b * p.vol + col, out + b * p.vol + col, m0, m1);
            }
}
template <class T>
static int t_axis0(int64_t batch, const int64_t* shape, int k0, int c0, const double* coef, int mode, int adjoint, const void* in, void* out, int chunk) {
    constexpr int VEC = 16 / (int)sizeof(T);
    Axis0P p;
    p.n0 = (int)shape[0]; p.plane = shape[1] * shape[2]; p.vol = (int64_t)p.n0 * p.plane;
    p.c0 = c0; p.mode = mode; p.adjoint = adjoint ? 1 : 0;
    p.pad_lo = adjoint ? k0 - 1 - c0 : c0; p.pad_hi = adjoint ? c0 : k0 - 1 - c0;
    p.lo_planes = p.hi_planes = 0;
    if (k0 < 2 || k0 > 9 || p.plane % VEC || c0 < 0 || c0 >= k0) return -103;
    for (int j = 0; j < 16; ++j) p.coef[j] = j < k0 ? coef[j] : 0.0;
    p.chunk = chunk > 0 ? std::min(chunk, p.n0) : p.n0;
    p.nchunk = (p.n0 + p.chunk - 1) / p.chunk;
    switch (k0) {
        case 2: t_axis0_k<T, 2>(p, batch, (const T*)in, (T*)out); break;
        case 3: t_axis0_k<T, 3>(p, batch, (const T*)in, (T*)out); break;
        case 4: t_axis0_k<T, 4>(p, batch, (const T*)in, (T*)out); break;
        case 5: t_axis0_k<T, 5>(p, batch, (const T*)in, (T*)out); break;
        case 6: t_axis0_k<T, 6>(p, batch, (const T*)in, (T*)out); break;
        case 7: t_axis0_k<T, 7>(p, batch, (const T*)in, (T*)out); break;
        case 8: t_axis0_k<T, 8>(p, batch, (const T*)in, (T*)out); break;
        case 9: t_axis0_k<T, 9>(p, batch, (const T*)in, (T*)out); break;
        default: return -103;
    }
    return 0;
}
extern "C" {
int emu_stencil_axis0_fold(int dtype, int64_t batch, const int64_t* shape, int k0, int c0, const double* coef, int mode, int adjoint, const void* in,
                           void* out, int chunk) {
    return dtype == PXB_F32 ? t_axis0<float>(batch, shape, k0, c0, coef, mode, adjoint, in, out, chunk)
                            : t_axis0<double>(batch, shape, k0, c0, coef, mode, adjoint, in, out, chunk);
}
int emu_stencil3d_dense(const pxb_stencil3d_dense* d, const void* in, void* out, int chunk) {
    g_d3_chunk = chunk;
    return d->dtype == PXB_F32 ? t_d3<float>(d, in, out) : t_d3<double>(d, in, out);
}
int emu_stencil3d(const pxb_stencil3d* d, const void* in, void* out) {
    return d->dtype == PXB_F32 ? t_st3<float>(d, in, out) : t_st3<double>(d, in, out);
}
int emu_tv_grad(int vec, int adjoint, const pxb_grad_desc* K, const void* in, void* out) {
    if (K->dtype == PXB_F32) return t_tvgrad_dir<float>(vec, adjoint != 0, K, in, out);
    return t_tvgrad_dir<double>(vec, adjoint != 0, K, in, out);
}
int emu_tv_tile2d(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, const void* u_in, const void* z_in, void* u_out, void* z_out,
                  void* x_out, double* nx, double* nz, int unused) {
    (void)unused;
    if (K->dtype == PXB_F32) return t_t2<float>(algo, K, p, u_in, z_in, u_out, z_out, x_out, nx, nz);
    return t_t2<double>(algo, K, p, u_in, z_in, u_out, z_out, x_out, nx, nz);
}
int emu_pad2d(const pxb_pad2d_desc* d, const void* in, void* ext) {
    if (d->dtype == PXB_F32) t_pad2d<float>(d, false, in, ext, 1.0, 0.0, nullptr, 0); else t_pad2d<double>(d, false, in, ext, 1.0, 0.0, nullptr, 0);
    return 0;
}
int emu_pad2d_adjoint(const pxb_pad2d_desc* d, const void* ext, void* out, double alpha, double beta, const void* add, int64_t add_period) {
    if (d->dtype == PXB_F32) t_pad2d<float>(d, true, ext, out, alpha, beta, add, add_period); else t_pad2d<double>(d, true, ext, out, alpha, beta, add, add_period);
    return 0;
}
int emu_stencil2d(const pxb_stencil2d* d, const void* in, void* out) {
    return d->dtype == PXB_F32 ? t_st2<float>(d, nullptr, 0, in, out) : t_st2<double>(d, nullptr, 0, in, out);
}
int emu_stencil2d_fista(const pxb_stencil2d* d, const pxb_fista_step* f, int which, void* out) {
    return d->dtype == PXB_F32 ? t_st2<float>(d, f, which, nullptr, out) : t_st2<double>(d, f, which, nullptr, out);
}
void emu_set_iter_modes(int on) { g_allow_modes = on != 0; }
long emu_w_global_cells(int reset) {
    const long v = g_w_global_cells;
    if (reset) g_w_global_cells = 0;
    return v;
}
int emu_tv_iter_tma(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, const void* u_in, const void* z_in, void* u_out, void* z_out,
                    void* x_out, double* nx, double* nz, int chunk) {
    if (K->dtype == PXB_F32) return t_tma<float>(algo, K, p, u_in, z_in, u_out, z_out, x_out, nx, nz, chunk);
    return t_tma<double>(algo, K, p, u_in, z_in, u_out, z_out, x_out, nx, nz, chunk);
}
int emu_set_edge_first(int v) {
    if (v != 1 && v != 2) return -1;
    g_emu_edge_first = v;
    return 0;
}
// the same with the halo exchange fused in (pxb_pds_iter_p2p): `peer` points into the neighbouring ranks' arrays of this process
int emu_tv_iter_tma_p2p(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, const void* u_in, const void* z_in, void* u_out, void* z_out,
                        void* x_out, double* nx, double* nz, int chunk, const pxb_peer* peer) {
    g_emu_peer = peer;
    const int rc = emu_tv_iter_tma(algo, K, p, u_in, z_in, u_out, z_out, x_out, nx, nz, chunk);
    g_emu_peer = nullptr;
    return rc;
}
int emu_tv_iter(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, const void* u_in, const void* z_in, void* u_out, void* z_out,
                void* x_out, double* nx, double* nz, int chunk) {
    if (K->dtype == PXB_F32) return t_iter<float>(algo, K, p, u_in, z_in, u_out, z_out, x_out, nx, nz, chunk);
    return t_iter<double>(algo, K, p, u_in, z_in, u_out, z_out, x_out, nx, nz, chunk);
}
// which: 0 primal, 1 dual.  Returns -2 when the descriptor is not eligible for the fast bodies.
int emu_tv_fast(int vec, int which, int algo, const pxb_grad_desc* K, const pxb_pds_params* p, void* xu, const void* z, void* x_out,
                void* w, double* norms) {
    if (K->dtype == PXB_F32) return t_tv_dir<float>(vec, which, algo, K, p, xu, z, x_out, w, norms);
    return t_tv_dir<double>(vec, which, algo, K, p, xu, z, x_out, w, norms);
}
int emu_stencil(const pxb_stencil_desc* d, int adjoint, const void* in, void* out) {
    if (d->dtype == PXB_F32) t_stencil<float>(d, adjoint, in, out); else t_stencil<double>(d, adjoint, in, out);
    return 0;
}
int emu_gradient(const pxb_grad_desc* d, int adjoint, const void* in, void* out) {
    if (d->dtype == PXB_F32) t_grad<float>(d, adjoint, in, out); else t_grad<double>(d, adjoint, in, out);
    return 0;
}
int emu_pds_primal(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, void* xu, const void* z, const void* ktz, void* x_out,
                   void* w, double* norms) {
    if (K->dtype == PXB_F32) t_primal<float>(algo, K, p, xu, z, ktz, x_out, w, norms);
    else t_primal<double>(algo, K, p, xu, z, ktz, x_out, w, norms);
    return 0;
}
int emu_pds_dual(const pxb_grad_desc* K, const pxb_pds_params* p, const void* w, void* z, double* norms) {
    if (K->dtype == PXB_F32) t_dual<float>(K, p, w, z, norms); else t_dual<double>(K, p, w, z, norms);
    return 0;
}
int emu_dual_update(int dtype, int kind, int64_t outer, int64_t group, int64_t inner, double lam, double sigma, double rho, void* z,
                    const void* t, double* norms) {
    for (int64_t o = 0; o < outer; ++o)
        for (int64_t i = 0; i < inner; ++i) {
            double a0 = 0, a1 = 0;
            if (dtype == PXB_F32) pxb_body_dual_update<float>(kind, group, inner, (float)lam, (float)sigma, (float)rho, (float*)z, (const float*)t, norms != nullptr, a0, a1, o, i);
            else pxb_body_dual_update<double>(kind, group, inner, lam, sigma, rho, (double*)z, (const double*)t, norms != nullptr, a0, a1, o, i);
            if (norms) { norms[2 * o] += a0; norms[2 * o + 1] += a1; }
        }
    return 0;
}
int emu_prox_l21(int dtype, int64_t outer, int64_t group, int64_t inner, double lam, double tau, const void* x, void* out) {
    for (int64_t o = 0; o < outer; ++o)
        for (int64_t i = 0; i < inner; ++i) {
            if (dtype == PXB_F32) pxb_body_prox_l21<float>(group, inner, (float)lam, (float)tau, (const float*)x, (float*)out, o, i);
            else pxb_body_prox_l21<double>(group, inner, lam, tau, (const double*)x, (double*)out, o, i);
        }
    return 0;
}
int emu_prox_lincomb(int dtype, const pxb_prox_spec* g, double tau, int64_t n, void* out, double a, const void* x, double b,
                     const void* y, int64_t ny, double c, const void* z, int64_t nz) {
    if (ny >= n) ny = 0;
    if (nz >= n) nz = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (dtype == PXB_F32) {
            float v = pxb_lincomb_at<float>((float)a, (const float*)x, (float)b, (const float*)y, ny, (float)c, (const float*)z, nz, i);
            if (g) v = pxb_prox_eval<float>(g->kind, (float)g->p0, (float)g->p1, v, (float)tau);
            ((float*)out)[i] = v;
        } else {
            double v = pxb_lincomb_at<double>(a, (const double*)x, b, (const double*)y, ny, c, (const double*)z, nz, i);
            if (g) v = pxb_prox_eval<double>(g->kind, g->p0, g->p1, v, tau);
            ((double*)out)[i] = v;
        }
    }
    return 0;
}
}
