// pxb_tv_tma.cu -- TMA + mbarrier pipelined single-kernel iteration for 3-D volumes (design: pxb_tv_tma.cuh).
//
//   thread 0 of each CTA:  cp.async.bulk.tensor.5d (SASS UTMALDG) x5 per plane into a 3-stage shared-memory ring,
//                          completion counted in bytes on one mbarrier per stage
//   all 256 threads:       wait(stage m) -> phase A out of shared memory -> __syncthreads -> [thread 0 refills the
//                          stage just consumed with plane m+3] -> phase C of plane m-1
// Shared memory per CTA (fp32): 3 x 28.5 KB stages + 21.8 KB w ring = 107 KB  -> 2 CTAs per SM.
#include "pxb_launch.cuh"
#include "pxb_tma_util.cuh"
#include "pxb_tv_tma.cuh"

namespace {

template <class T, int VEC, int TY, int ALGO, bool NORMS, class S, bool MODES = false>
__global__ void __launch_bounds__(32 * TY, 2)
    k_tv_iter_tma(const __grid_constant__ PxbTvP<T> q, const __grid_constant__ PxbIterGeom g, const __grid_constant__ PxbTmaGeom tg,
                  const __grid_constant__ PxbIterPtr<T> a, const __grid_constant__ CUtensorMap map_u, const __grid_constant__ CUtensorMap map_s,
                  const __grid_constant__ CUtensorMap map_z, const __grid_constant__ CUtensorMap map_z1) {
    using C = PxbTmaCfg<T, VEC, TY>;
    if (NORMS && pxb_iter_stopped(a.stop)) return;  // an earlier iteration of this batch met the stopping rule
    extern __shared__ __align__(128) unsigned char pxb_tma_smem[];
    T* stages = reinterpret_cast<T*>(pxb_tma_smem);
    T* ring = reinterpret_cast<T*>(pxb_tma_smem + C::SMEM_STAGES);
    uint64_t* full = reinterpret_cast<uint64_t*>(pxb_tma_smem + C::SMEM_STAGES + C::SMEM_RING);

    PxbIterItem it = pxb_iter_item(g, (int64_t)blockIdx.x, TY, C::T2);
    it.nopeer = (a.peer.dn_u == nullptr && a.peer.dn_z == nullptr && a.peer.up_z0 == nullptr) ? 1 : 0;
    const PxbIterRange R = pxb_iter_range<T>(q, it);
    const int tid = threadIdx.x;
    const bool need_next = pxb_has_cm<S>(q, 0);          // phase A of plane m reads z0 of plane m+1
    const int mload_hi = R.mhi + (need_next ? 1 : 0);    // planes [mlo, mload_hi) are staged
    const int b = (int)it.b;

    // one plane -> one stage: 4 or 5 boxes, all signalling the stage's mbarrier
    auto issue = [&](int m) {
        const int s = (m - R.mlo) % C::NSTAGE;
        T* st = stages + s * C::STAGE;
        uint64_t* bar = full + s;
        const bool staged_shift = S::FK >= 1 || tg.has_shift;
        const uint32_t bytes = C::BYTES_BOX * (3 + (staged_shift ? 1 : 0)) + C::BYTES_BOX1;
        mbar_expect_tx(bar, bytes);
        const int cc = it.c0 - VEC, cr = it.r0 - 1, cp = m + tg.gl;
        tma_load_5d(st + C::OFF_U, &map_u, bar, cc, cr, cp, 0, b);
        if (staged_shift) tma_load_5d(st + C::OFF_S, &map_s, bar, cc, cr, cp, 0, tg.sh_batched ? b : 0);
        tma_load_5d(st + C::OFF_Z0, &map_z, bar, cc, cr, cp, 0, b);
        tma_load_5d(st + C::OFF_Z2, &map_z, bar, cc, cr, cp, 2, b);
        tma_load_5d(st + C::OFF_Z1, &map_z1, bar, cc, cr - 1, cp, 1, b);
    };

    if (tid == 0) {
        for (int s = 0; s < C::NSTAGE; ++s) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        // peer-memory exchange: the ghost planes this work item reads were stored by the neighbour's previous iteration
        if (a.peer.lo_wait != nullptr && it.m0 == 0) pxb_peer_wait(a.peer.lo_wait, a.peer.target);
        if (a.peer.hi_wait != nullptr && it.m1 == g.nM) pxb_peer_wait(a.peer.hi_wait, a.peer.target);
        if (a.peer.lo_wait != nullptr || a.peer.hi_wait != nullptr) asm volatile("fence.proxy.async;" ::: "memory");  // ... and are read by TMA
    }
    __syncthreads();
    if (tid == 0)
        for (int m = R.mlo; m < R.mlo + C::NSTAGE && m < mload_hi; ++m) issue(m);

    PxbTmaThread<T, VEC> th;
    for (int k = 0; k < 3; ++k)
        for (int j = 0; j < VEC; ++j) th.zc[k][j] = th.zprev[k][j] = T(0);
    for (int k = 0; k < 4; ++k) th.acc[k] = 0.0;
    pxb_tma_prologue<T, VEC, TY, MODES, ALGO == PXB_PD3O && NORMS>(q, g, it, a, tid, R.mlo, th);

    const int lag = S::SCHEME == PXB_SCHEME_FWD ? 1 : R.lag;
    const PxbTmaFold fb = MODES ? pxb_tma_fold_setup<T, VEC, TY>(q, g, it) : PxbTmaFold{};
    int s = 0;            // stage of plane m, and the parity of its mbarrier phase
    uint32_t par = 0;
    for (int m = R.mlo; m < R.mhi; ++m) {
        mbar_wait(full + s, par);
        const T* st = stages + s * C::STAGE;
        const T* st_next = st;
        const int s1 = s + 1 == C::NSTAGE ? 0 : s + 1;
        if (need_next) {
            mbar_wait(full + s1, s1 == 0 ? par ^ 1u : par);
            st_next = stages + s1 * C::STAGE;
        }
        pxb_tma_phaseA<T, VEC, TY, ALGO, NORMS, S, MODES>(q, g, tg, it, a, tid, m, st, st_next, ring, th, fb);
        __syncthreads();  // w(m) complete; every thread is done with stage s
        if (tid == 0 && m + C::NSTAGE < mload_hi) issue(m + C::NSTAGE);
        const int mm = m - lag;
        if (mm >= it.m0 && mm < it.m1) {
            T zo[3][VEC];
            for (int kk = 0; kk < 3; ++kk)
                for (int j = 0; j < VEC; ++j) zo[kk][j] = lag ? th.zprev[kk][j] : th.zc[kk][j];
            pxb_iter_phaseC<T, VEC, C::TXL, TY, 3, NORMS, S>(q, g, it, a, tid, mm, ring, zo, th.acc, th.pz);
            th.pz += g.sM;
        }
        for (int kk = 0; kk < 3; ++kk)
            for (int j = 0; j < VEC; ++j) th.zprev[kk][j] = th.zc[kk][j];
        s = s1;
        if (s == 0) par ^= 1u;
    }
    // a stage filled for `need_next` beyond the last plane has been waited on above (k1), nothing is in flight here

    {   // peer-memory exchange: this work item's share of the boundary planes is in the neighbours' memory -> tell them
        const bool dn = a.peer.dn_flag != nullptr && it.m0 == 0, up = a.peer.up_flag != nullptr && it.m1 == g.nM;
        if (dn || up) {
            __threadfence_system();
            __syncthreads();
            if (tid == 0) {
                if (dn) atomicAdd_system(a.peer.dn_flag, 1u);
                if (up) atomicAdd_system(a.peer.up_flag, 1u);
            }
        }
    }

    if (NORMS && !PXB_EXP(2)) {
        __shared__ double red[4][C::NT / 32];
        double v[4] = {th.acc[0], th.acc[1], th.acc[2], th.acc[3]};
        for (int o = 16; o > 0; o >>= 1)
            for (int kk = 0; kk < 4; ++kk) v[kk] += __shfl_down_sync(0xffffffffu, v[kk], o);
        const int w = tid >> 5, l = tid & 31;
        if (l == 0)
            for (int kk = 0; kk < 4; ++kk) red[kk][w] = v[kk];
        __syncthreads();
        if (tid == 0) {
            double s4[4] = {0.0, 0.0, 0.0, 0.0};
            for (int i = 0; i < C::NT / 32; ++i)
                for (int kk = 0; kk < 4; ++kk) s4[kk] += red[kk][i];
            if (a.norms_x && !PXB_EXP(1)) { atomicAdd(a.norms_x + 2 * it.b, s4[0]); atomicAdd(a.norms_x + 2 * it.b + 1, s4[1]); }
            if (a.norms_z && !PXB_EXP(1)) { atomicAdd(a.norms_z + 2 * it.b, s4[2]); atomicAdd(a.norms_z + 2 * it.b + 1, s4[3]); }
            pxb_iter_finish(a.stop, a.norms_x, a.norms_z, gridDim.x);
        }
    }
}

template <class T, int ALGO, bool NORMS>
int run(const pxb_grad_desc& d, const pxb_pds_params& P, const PxbIterPtr<T>& a_in, int chunk_hint, cudaStream_t s, cudaError_t* err, const pxb_peer* peer) {
    constexpr int VEC = 16 / (int)sizeof(T), TY = 8;
    using C = PxbTmaCfg<T, VEC, TY>;
    PxbTvCoef cf;
    PxbIterGeom g;
    PxbTvP<T> q;
    if (int why = pxb_iter_setup(d, P, VEC, TY, C::T2, chunk_hint, 148 * 2, cf, g, pxb_iter_modes() != 0)) return why;
    PxbIterPtr<T> a = a_in;
    if (peer) {  // halo exchange through peer memory: edge chunks first, counters in units of one per edge thread block
        if (d.batch != 1) return 24;
        // (PXB_P2P_INTERLEAVE=1: the edge work items one in four instead of all first -- measured on 2 B200s: no difference, 3.21 ms either way)
        static const int interleave = [] { const char* e = getenv("PXB_P2P_INTERLEAVE"); return (e && e[0] == '1') ? 2 : 1; }();
        g.edge_first = interleave;
        a.peer.dn_u = (T*)peer->dn_u; a.peer.dn_z = (T*)peer->dn_z; a.peer.dn_zvol = peer->dn_zvol; a.peer.up_z0 = (T*)peer->up_z0;
        a.peer.dn_flag = peer->dn_flag; a.peer.up_flag = peer->up_flag; a.peer.lo_wait = peer->lo_wait; a.peer.hi_wait = peer->hi_wait;
        a.peer.target = (unsigned)((uint64_t)peer->epoch * (uint64_t)g.ntR * (uint64_t)g.ntC);
    }
    pxb_tv_prepare<T>(d, cf, P, q);
    PxbTmaGeom tg;
    PxbTmaBoxDesc mu, ms, mz;
    if (int why = pxb_tma_setup<T, VEC, TY>(d, P, cf, g, q, a.u_in, a.z_in, tg, mu, ms, mz)) return why;
    PxbTmaBoxDesc mz1 = mz;
    mz1.box[1] = C::BR1;
    alignas(64) CUtensorMap tu, ts, tz, tz1;
    auto enc = [](const PxbTmaBoxDesc& m, CUtensorMap* out) { return pxb_tma_encode_cached<T>(5, m.base, m.dim, m.stride, m.box, out); };
    if (!enc(mu, &tu) || !enc(ms, &ts) || !enc(mz, &tz) || !enc(mz1, &tz1)) return 23;
    // specialised instances: forward differences + L21 + shifted squared-l2 data term staged per voxel + g in
    // {positivity, none}; everything else runs the generic instance
    const int spec = pxb_tma_pick_spec<T>(cf, q, tg);
#ifdef PXB_EXPERIMENT
    {
        static int exp_set = -1;
        const char* e = getenv("PXB_EXP");
        const int v = e ? atoi(e) : 0;
        if (v != exp_set) { cudaMemcpyToSymbol(pxb_exp_flags, &v, sizeof(int)); exp_set = v; }
    }
#endif
    auto go = [&](auto kern) {
        // (the attribute is per function: set it on every launch path once; cheap enough to repeat)
        cudaError_t e = pxb_smem_attr_once((const void*)kern, (int)C::SMEM);
        if (e != cudaSuccess) { *err = e; return; }
        kern<<<(unsigned)g.nblocks, C::NT, C::SMEM, s>>>(q, g, tg, a, tu, ts, tz, tz1);
        *err = cudaGetLastError();
    };
    if (pxb_any_mode(d)) {  // folding boundary modes: MODES instances (fold terms of K^T z, folded rim of the w ring)
        if (spec == 1) go(k_tv_iter_tma<T, VEC, TY, ALGO, NORMS, PxbSpecFwdPos, true>);
        else if (spec == 2) go(k_tv_iter_tma<T, VEC, TY, ALGO, NORMS, PxbSpecFwdNone, true>);
        else go(k_tv_iter_tma<T, VEC, TY, ALGO, NORMS, PxbSpecAny, true>);
        return 0;
    }
    if (spec == 1) go(k_tv_iter_tma<T, VEC, TY, ALGO, NORMS, PxbSpecFwdPos>);
    else if (spec == 2) go(k_tv_iter_tma<T, VEC, TY, ALGO, NORMS, PxbSpecFwdNone>);
    else if (spec == 3 && ALGO == PXB_CV) go(k_tv_iter_tma<T, VEC, TY, PXB_CV, NORMS, PxbSpecFwdPosG>);
    else if (spec == 4 && ALGO == PXB_CV) go(k_tv_iter_tma<T, VEC, TY, PXB_CV, NORMS, PxbSpecFwdNoneG>);
    else go(k_tv_iter_tma<T, VEC, TY, ALGO, NORMS, PxbSpecAny>);
    return 0;
}

template <class T>
int dispatch(int algo, const pxb_grad_desc& d, const pxb_pds_params& P, const PxbIterPtr<T>& a, int chunk_hint, cudaStream_t s, cudaError_t* err,
             const pxb_peer* peer) {
    const bool norms = a.norms_x || a.norms_z;
    if (algo == PXB_PD3O) return norms ? run<T, PXB_PD3O, true>(d, P, a, chunk_hint, s, err, peer) : run<T, PXB_PD3O, false>(d, P, a, chunk_hint, s, err, peer);
    return norms ? run<T, PXB_CV, true>(d, P, a, chunk_hint, s, err, peer) : run<T, PXB_CV, false>(d, P, a, chunk_hint, s, err, peer);
}

}  // namespace

// > 0: not eligible (reason), 0: launched (or *err set)
int pxb_tv_tma_try(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, const void* xu_in, const void* z_in, void* xu_out, void* z_out,
                   void* x_out, double* norms_x, double* norms_z, int chunk_hint, cudaStream_t s, cudaError_t* err, const PxbIterStop* stop,
                   const pxb_peer* peer) {
    if (K->ndir != 3) return 20;
    const PxbIterStop st = stop ? *stop : PxbIterStop{};
    if (K->dtype == PXB_F32) {
        PxbIterPtr<float> a{(const float*)xu_in, (const float*)z_in, (float*)xu_out, (float*)z_out, (float*)x_out, norms_x, norms_z, st, {}};
        return dispatch<float>(algo, *K, *p, a, chunk_hint, s, err, peer);
    }
    PxbIterPtr<double> a{(const double*)xu_in, (const double*)z_in, (double*)xu_out, (double*)z_out, (double*)x_out, norms_x, norms_z, st, {}};
    return dispatch<double>(algo, *K, *p, a, chunk_hint, s, err, peer);
}
