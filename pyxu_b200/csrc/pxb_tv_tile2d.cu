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

// ---------------------------------------------------------------------------------------------------------
// n iterations in ONE launch for images small enough that every tile has its own resident CTA (512^2 fp64 = 256 CTAs of the 592
// the device holds): a problem of that size is launch-bound -- 2 MiB per field, the state lives in L2, one iteration is ~3 us of
// work against ~11 us launch to launch.  Each CTA keeps its tile; between iterations a grid-wide barrier (one counter, arrive +
// spin with acquire, bounded, then trap); even iterations read pair A and write pair B, odd ones the reverse (two sets of tensor
// maps).  The stopping rule is evaluated exactly as in the one-launch-per-iteration form (pxb_iter_finish by the block that draws
// the last ticket, BEFORE it arrives at the barrier), every block reads the flag after the barrier.  Launched cooperatively: the
// runtime refuses the launch if the grid cannot be co-resident.
// ---------------------------------------------------------------------------------------------------------
struct PxbT2Loop {
    int n;               // iterations
    int64_t per;         // doubles between the sums of consecutive iterations (0: no sums)
    int nx_on, nz_on;    // which sums are accumulated (offsets 0 and 2 * batch inside an iteration's block)
    int64_t batch;
    unsigned* gbar;      // grid barrier counter, zero at launch
};

static __device__ __forceinline__ void pxb_grid_arrive_wait(unsigned* gbar, unsigned target) {
    __threadfence();
    atomicAdd(gbar, 1u);
    for (unsigned spin = 0; spin < (1u << 26); ++spin) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gbar) : "memory");
        if ((int)(v - target) >= 0) return;
    }
    __trap();
}

template <class T, int VEC, int ALGO, bool NORMS, class S>
__global__ void __launch_bounds__(256, 4)
    k_tv_tile2d_loop(const __grid_constant__ PxbTvP<T> q, const __grid_constant__ PxbT2Geom g, const __grid_constant__ PxbIterPtr<T> a0,
                     const __grid_constant__ PxbIterPtr<T> a1, const __grid_constant__ PxbT2Loop L, const __grid_constant__ CUtensorMap map_s,
                     const __grid_constant__ CUtensorMap ua, const __grid_constant__ CUtensorMap zra, const __grid_constant__ CUtensorMap zca,
                     const __grid_constant__ CUtensorMap ub, const __grid_constant__ CUtensorMap zrb, const __grid_constant__ CUtensorMap zcb) {
    using C = PxbT2Cfg<T, VEC>;
    extern __shared__ __align__(128) unsigned char pxb_t2_smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ int stop_now;
    __shared__ double red[4][8];
    T* sm = reinterpret_cast<T*>(pxb_t2_smem);
    const PxbT2Item it = pxb_t2_item(g, (int64_t)blockIdx.x, C::TY, C::T2);
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    for (int i = 0; i < L.n; ++i) {
        const bool even = (i & 1) == 0;
        PxbIterPtr<T> a = even ? a0 : a1;
        if (NORMS) {
            double* blk = a0.norms_x ? a0.norms_x : a0.norms_z;  // (launcher: both point at the first iteration's block)
            a.norms_x = L.nx_on ? blk + (int64_t)i * L.per : nullptr;
            a.norms_z = L.nz_on ? blk + (int64_t)i * L.per + 2 * L.batch : nullptr;
        }
        if (tid == 0) {
            mbar_expect_tx(&bar, C::BYTES_BOX * (2 + (g.has_shift ? 1 : 0)) + C::BYTES_BOXZ);
            const int cc = it.c0 - VEC, cr = it.r0 - 1;
            tma_load_3d(sm + C::OFF_U, even ? &ua : &ub, &bar, cc, cr, (int)it.img);
            if (g.has_shift) tma_load_3d(sm + C::OFF_S, &map_s, &bar, cc, cr, g.sh_mode ? it.i0 : (int)it.img);
            const int pz = (int)(it.b * 2 * g.n0) + it.i0;
            tma_load_3d(sm + C::OFF_ZR, even ? &zra : &zrb, &bar, cc, cr - 1, pz);
            tma_load_3d(sm + C::OFF_ZC, even ? &zca : &zcb, &bar, cc, cr, pz + g.n0);
        }
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        mbar_wait(&bar, (uint32_t)(i & 1));
        pxb_t2_phaseA<T, VEC, ALGO, NORMS, S, false>(q, g, it, a, tid, sm, acc);
        __syncthreads();
        pxb_t2_phaseC<T, VEC, NORMS, S>(q, g, it, a, tid, sm, acc);
        if (NORMS) {
            for (int o = 16; o > 0; o >>= 1)
                for (int k = 0; k < 4; ++k) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], o);
            if ((tid & 31) == 0)
                for (int k = 0; k < 4; ++k) red[k][tid >> 5] = acc[k];
        }
        __syncthreads();  // this block's stores of the iteration are issued; the staged boxes are free
        if (tid == 0) {
            if (NORMS) {
                double s4[4] = {0.0, 0.0, 0.0, 0.0};
                for (int w = 0; w < 8; ++w)
                    for (int k = 0; k < 4; ++k) s4[k] += red[k][w];
                if (a.norms_x) { atomicAdd(a.norms_x + 2 * it.b, s4[0]); atomicAdd(a.norms_x + 2 * it.b + 1, s4[1]); }
                if (a.norms_z) { atomicAdd(a.norms_z + 2 * it.b, s4[2]); atomicAdd(a.norms_z + 2 * it.b + 1, s4[3]); }
                pxb_iter_finish(a.stop, a.norms_x, a.norms_z, gridDim.x);
            }
            pxb_grid_arrive_wait(L.gbar, (unsigned)(i + 1) * gridDim.x);
            stop_now = (NORMS && pxb_iter_stopped(a.stop)) ? 1 : 0;
            asm volatile("fence.proxy.async;" ::: "memory");  // the other blocks' stores are read by this block's TMA loads next
        }
        __syncthreads();
        if (stop_now) return;
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

unsigned* grid_barrier_word(cudaError_t* err) {  // one zero-initialised word per device, reset on the stream before every use
    static unsigned* word[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return nullptr;
    if (!word[dev]) {
        *err = cudaMalloc((void**)&word[dev], 256);
        if (*err != cudaSuccess) return nullptr;
    }
    return word[dev];
}

template <class T, int ALGO, bool NORMS>
int run_loop(const pxb_grad_desc& d, const pxb_pds_params& P, const PxbIterPtr<T>& a0, const PxbIterPtr<T>& a1, PxbT2Loop L, cudaStream_t s, cudaError_t* err) {
    constexpr int VEC = 16 / (int)sizeof(T);
    using C = PxbT2Cfg<T, VEC>;
    PxbTvCoef cf;
    PxbTvP<T> q;
    PxbT2Geom g;
    if (int why = pxb_t2_setup<T, VEC>(d, P, cf, q, g, pxb_iter_modes() != 0)) return why;
    alignas(64) CUtensorMap ts, m[2][3];
    const uint64_t stride[3] = {1, (uint64_t)g.n2, (uint64_t)g.s0};
    const uint64_t dim_u[3] = {(uint64_t)g.n2, (uint64_t)g.n1, (uint64_t)g.nimg};
    const uint64_t dim_z[3] = {(uint64_t)g.n2, (uint64_t)g.n1, (uint64_t)g.nimg * 2};
    const uint64_t dim_s[3] = {(uint64_t)g.n2, (uint64_t)g.n1, (uint64_t)(g.sh_mode ? g.n0 : g.nimg)};
    const uint32_t box[3] = {(uint32_t)C::BW, (uint32_t)C::BR, 1}, boxz[3] = {(uint32_t)C::BW, (uint32_t)C::BRZ, 1};
    const void* sptr = q.fkind == PXB_F_GRADARR ? (const void*)q.garr : (const void*)q.shift;
    if (!pxb_tma_encode_cached<T>(3, g.has_shift ? sptr : (const void*)a0.u_in, g.has_shift ? dim_s : dim_u, stride, box, &ts)) return 23;
    const PxbIterPtr<T>* pair[2] = {&a0, &a1};
    for (int k = 0; k < 2; ++k)
        if (!pxb_tma_encode_cached<T>(3, pair[k]->u_in, dim_u, stride, box, &m[k][0]) || !pxb_tma_encode_cached<T>(3, pair[k]->z_in, dim_z, stride, boxz, &m[k][1]) ||
            !pxb_tma_encode_cached<T>(3, pair[k]->z_in, dim_z, stride, box, &m[k][2]))
            return 23;
    bool fwd = true;
    for (int k = 0; k < 2; ++k) fwd = fwd && cf.cm[k] == 0.0 && cf.cp[k] != 0.0;
    L.gbar = grid_barrier_word(err);
    if (!L.gbar) return *err != cudaSuccess ? 0 : 24;
    int rc = 0;
    auto go = [&](auto kern) {
        cudaError_t e = pxb_smem_attr_once((const void*)kern, (int)C::SMEM);
        if (e != cudaSuccess) { *err = e; return; }
        int per_sm = 0, dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, C::NT, C::SMEM) != cudaSuccess || (int64_t)per_sm * sms < g.nblocks) {
            rc = 25;  // the tiles do not all fit on the device at once: one launch per iteration
            return;
        }
        e = cudaMemsetAsync(L.gbar, 0, sizeof(unsigned), s);
        if (e != cudaSuccess) { *err = e; return; }
        void* args[] = {(void*)&q, (void*)&g, (void*)&a0, (void*)&a1, (void*)&L, (void*)&ts, (void*)&m[0][0], (void*)&m[0][1], (void*)&m[0][2],
                        (void*)&m[1][0], (void*)&m[1][1], (void*)&m[1][2]};
        *err = cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)g.nblocks), dim3(C::NT), args, C::SMEM, s);
    };
    // Folding boundary modes stay with one launch per iteration: their fold terms / folded rims read the other blocks' results with
    // ordinary (L1-cached, possibly non-coherent) loads, which only a kernel boundary makes safe; here every cross-block operand
    // comes through TMA (L2) after the barrier's acquire + proxy fence.
    if (pxb_any_mode(d)) return 26;
    if (fwd && q.hkind == PXB_DUAL_L21) go(k_tv_tile2d_loop<T, VEC, ALGO, NORMS, PxbSpec<PXB_SCHEME_FWD, -1, PXB_DUAL_L21, -1>>);
    else go(k_tv_tile2d_loop<T, VEC, ALGO, NORMS, PxbSpecAny>);
    return rc;
}

}  // namespace

// n iterations in one cooperative launch (k_tv_tile2d_loop); > 0: not eligible (reason; nothing launched), 0: launched (or *err set).
// norms: block of the first iteration (sums of iteration i at norms + i * per), use_x / use_z: which sums the rule reads.
int pxb_tv_tile2d_loop_try(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, void* xu_a, void* z_a, void* xu_b, void* z_b, void* x_out, double* norms,
                           int use_x, int use_z, int n, cudaStream_t s, cudaError_t* err, const PxbIterStop* stop) {
    if (K->ndir != 2) return 2;
    const bool sums = norms && (use_x || use_z);
    const PxbIterStop st = stop ? *stop : PxbIterStop{};
    PxbT2Loop L{n, sums ? 4 * K->batch : 0, sums && use_x ? 1 : 0, sums && use_z ? 1 : 0, K->batch, nullptr};
#define PXB_T2_LOOP(T)                                                                                                          \
    {                                                                                                                           \
        PxbIterPtr<T> a0{(const T*)xu_a, (const T*)z_a, (T*)xu_b, (T*)z_b, (T*)x_out, sums ? norms : nullptr, sums ? norms : nullptr, st}; \
        PxbIterPtr<T> a1{(const T*)xu_b, (const T*)z_b, (T*)xu_a, (T*)z_a, (T*)x_out, sums ? norms : nullptr, sums ? norms : nullptr, st}; \
        if (algo == PXB_PD3O) return sums ? run_loop<T, PXB_PD3O, true>(*K, *p, a0, a1, L, s, err) : run_loop<T, PXB_PD3O, false>(*K, *p, a0, a1, L, s, err); \
        return sums ? run_loop<T, PXB_CV, true>(*K, *p, a0, a1, L, s, err) : run_loop<T, PXB_CV, false>(*K, *p, a0, a1, L, s, err); \
    }
    if (K->dtype == PXB_F32) PXB_T2_LOOP(float)
    PXB_T2_LOOP(double)
#undef PXB_T2_LOOP
}

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
