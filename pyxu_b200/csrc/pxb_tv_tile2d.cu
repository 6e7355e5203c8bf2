// pxb_tv_tile2d.cu -- launcher of the TMA-tiled single-kernel iteration for 2-D TV problems (design: pxb_tv_tile2d.cuh).
#include "pxb_launch.cuh"
#include "pxb_tma_util.cuh"
#include "pxb_tv_tile2d.cuh"

namespace {

template <class T, int VEC, int ALGO, bool NORMS, class S, bool MODES = false>
__global__ void __launch_bounds__(256, 4)
    k_tv_tile2d(const __grid_constant__ PxbTvP<T> q, const __grid_constant__ PxbT2Geom g, const __grid_constant__ PxbIterPtr<T> a,
                const __grid_constant__ CUtensorMap map_u, const __grid_constant__ CUtensorMap map_s, const __grid_constant__ CUtensorMap map_zr,
                const __grid_constant__ CUtensorMap map_zc) {
    using C = PxbT2Cfg<T, VEC>;
    if (NORMS && pxb_iter_stopped(a.stop)) return;  // an earlier iteration of this batch met the stopping rule
    extern __shared__ __align__(128) unsigned char pxb_t2_smem[];
    __shared__ __align__(8) uint64_t bar;
    T* sm = reinterpret_cast<T*>(pxb_t2_smem);
    const PxbT2Item it = pxb_t2_item(g, (int64_t)blockIdx.x, C::TY, C::T2);
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, C::BYTES_BOX * (2 + (g.has_shift ? 1 : 0)) + C::BYTES_BOXZ);
        const int cc = it.c0 - VEC, cr = it.r0 - 1;
        tma_load_3d(sm + C::OFF_U, &map_u, &bar, cc, cr, (int)it.img);
        if (g.has_shift) tma_load_3d(sm + C::OFF_S, &map_s, &bar, cc, cr, g.sh_mode ? it.i0 : (int)it.img);
        const int pz = (int)(it.b * 2 * g.n0) + it.i0;
        tma_load_3d(sm + C::OFF_ZR, &map_zr, &bar, cc, cr - 1, pz);
        tma_load_3d(sm + C::OFF_ZC, &map_zc, &bar, cc, cr, pz + g.n0);
    }
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    mbar_wait(&bar, 0);
    pxb_t2_phaseA<T, VEC, ALGO, NORMS, S, MODES>(q, g, it, a, tid, sm, acc);
    __syncthreads();
    pxb_t2_phaseC<T, VEC, NORMS, S>(q, g, it, a, tid, sm, acc);
    if (NORMS) {
        __shared__ double red[4][8];
        for (int o = 16; o > 0; o >>= 1)
            for (int k = 0; k < 4; ++k) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], o);
        if ((tid & 31) == 0)
            for (int k = 0; k < 4; ++k) red[k][tid >> 5] = acc[k];
        __syncthreads();
        if (tid == 0) {
            double s4[4] = {0.0, 0.0, 0.0, 0.0};
            for (int i = 0; i < 8; ++i)
                for (int k = 0; k < 4; ++k) s4[k] += red[k][i];
            if (a.norms_x) { atomicAdd(a.norms_x + 2 * it.b, s4[0]); atomicAdd(a.norms_x + 2 * it.b + 1, s4[1]); }
            if (a.norms_z) { atomicAdd(a.norms_z + 2 * it.b, s4[2]); atomicAdd(a.norms_z + 2 * it.b + 1, s4[3]); }
            pxb_iter_finish(a.stop, a.norms_x, a.norms_z, gridDim.x);
        }
    }
}

template <class T, int ALGO, bool NORMS>
int run(const pxb_grad_desc& d, const pxb_pds_params& P, const PxbIterPtr<T>& a, cudaStream_t s, cudaError_t* err) {
    constexpr int VEC = 16 / (int)sizeof(T);
    using C = PxbT2Cfg<T, VEC>;
    PxbTvCoef cf;
    PxbTvP<T> q;
    PxbT2Geom g;
    if (int why = pxb_t2_setup<T, VEC>(d, P, cf, q, g, pxb_iter_modes() != 0)) return why;
    alignas(64) CUtensorMap tu, ts, tzr, tzc;
    const uint64_t stride[3] = {1, (uint64_t)g.n2, (uint64_t)g.s0};
    const uint64_t dim_u[3] = {(uint64_t)g.n2, (uint64_t)g.n1, (uint64_t)g.nimg};
    const uint64_t dim_z[3] = {(uint64_t)g.n2, (uint64_t)g.n1, (uint64_t)g.nimg * 2};
    const uint64_t dim_s[3] = {(uint64_t)g.n2, (uint64_t)g.n1, (uint64_t)(g.sh_mode ? g.n0 : g.nimg)};
    const uint32_t box[3] = {(uint32_t)C::BW, (uint32_t)C::BR, 1}, boxz[3] = {(uint32_t)C::BW, (uint32_t)C::BRZ, 1};
    const void* sptr = q.fkind == PXB_F_GRADARR ? (const void*)q.garr : (const void*)q.shift;
    if (!pxb_tma_encode_cached<T>(3, a.u_in, dim_u, stride, box, &tu) || !pxb_tma_encode_cached<T>(3, g.has_shift ? sptr : (const void*)a.u_in, g.has_shift ? dim_s : dim_u, stride, box, &ts) ||
        !pxb_tma_encode_cached<T>(3, a.z_in, dim_z, stride, boxz, &tzr) || !pxb_tma_encode_cached<T>(3, a.z_in, dim_z, stride, box, &tzc))
        return 23;
    bool fwd = true;
    for (int k = 0; k < 2; ++k) fwd = fwd && cf.cm[k] == 0.0 && cf.cp[k] != 0.0;
    auto go = [&](auto kern) {
        cudaError_t e = pxb_smem_attr_once((const void*)kern, (int)C::SMEM);
        if (e != cudaSuccess) { *err = e; return; }
        kern<<<(unsigned)g.nblocks, C::NT, C::SMEM, s>>>(q, g, a, tu, ts, tzr, tzc);
        *err = cudaGetLastError();
    };
    if (pxb_any_mode(d)) {  // folding boundary modes: MODES instances (fold terms of K^T z, folded rim of the w tile)
        if (fwd && q.hkind == PXB_DUAL_L21) go(k_tv_tile2d<T, VEC, ALGO, NORMS, PxbSpec<PXB_SCHEME_FWD, -1, PXB_DUAL_L21, -1>, true>);
        else go(k_tv_tile2d<T, VEC, ALGO, NORMS, PxbSpecAny, true>);
        return 0;
    }
    if (fwd && q.hkind == PXB_DUAL_L21) go(k_tv_tile2d<T, VEC, ALGO, NORMS, PxbSpec<PXB_SCHEME_FWD, -1, PXB_DUAL_L21, -1>>);
    else go(k_tv_tile2d<T, VEC, ALGO, NORMS, PxbSpecAny>);
    return 0;
}

}  // namespace

// > 0: not eligible (reason), 0: launched (or *err set)
int pxb_tv_tile2d_try(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, const void* xu_in, const void* z_in, void* xu_out, void* z_out,
                      void* x_out, double* norms_x, double* norms_z, cudaStream_t s, cudaError_t* err, const PxbIterStop* stop) {
    if (K->ndir != 2) return 2;
    const bool norms = norms_x || norms_z;
    const PxbIterStop st = stop ? *stop : PxbIterStop{};
#define PXB_T2_GO(T)                                                                                                  \
    {                                                                                                                 \
        PxbIterPtr<T> a{(const T*)xu_in, (const T*)z_in, (T*)xu_out, (T*)z_out, (T*)x_out, norms_x, norms_z, st};     \
        if (algo == PXB_PD3O) return norms ? run<T, PXB_PD3O, true>(*K, *p, a, s, err) : run<T, PXB_PD3O, false>(*K, *p, a, s, err); \
        return norms ? run<T, PXB_CV, true>(*K, *p, a, s, err) : run<T, PXB_CV, false>(*K, *p, a, s, err);            \
    }
    if (K->dtype == PXB_F32) PXB_T2_GO(float)
    PXB_T2_GO(double)
#undef PXB_T2_GO
}
