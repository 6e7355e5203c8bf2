// pxb_stencil3d.cu -- launcher of the single-pass separable 3-D stencil (design: pxb_stencil3d.cuh).
#include "pxb_launch.cuh"
#include "pxb_tma_util.cuh"
#include "pxb_stencil3d.cuh"

int pxb_st3_fast_try(int dtype, const PxbSt3P& g, const void* in, void* out, cudaStream_t s, cudaError_t* err);  // pxb_stencil3d_fast.cu

namespace {

int g_st3_path = 0;  // pxb_set_stencil3d_path: 0 = the K x K x K instances where they apply, 1 = k_stencil3d always

template <class T, int VEC, int NV, int K0>
__global__ void __launch_bounds__(256) k_stencil3d(const __grid_constant__ PxbSt3P p, const __grid_constant__ CUtensorMap map, T* __restrict__ out) {
    using C = PxbSt3Cfg<T, VEC>;
    extern __shared__ __align__(128) unsigned char pxb_st3_smem[];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ T c1s[PXB_ST2_MAXTAP + 2 * (C::R - 1)];
    T* stage0 = reinterpret_cast<T*>(pxb_st3_smem);
    const int box_elems = (p.s.bh * p.s.bw + 31) / 32 * 32;
    T* mid = stage0 + 2 * box_elems;
    const int tid = threadIdx.x;
    unsigned blk = blockIdx.x;
    const int tx = blk % (unsigned)p.s.ntx; blk /= (unsigned)p.s.ntx;
    const int ty = blk % (unsigned)p.s.nty; blk /= (unsigned)p.s.nty;
    const int ch = blk % (unsigned)p.nchunk;
    const int b = blk / (unsigned)p.nchunk;
    const int x0 = tx * C::TX, y0 = ty * C::TY;
    const int m0 = ch * p.chunk, m1 = min(p.n0, m0 + p.chunk);
    const int pl_lo = m0 - p.c0, pl_hi = m1 + K0 - 1 - p.c0;            // input planes this chunk needs
    const int ra = max(pl_lo, -p.lo_planes), rb = min(pl_hi, p.n0 + p.hi_planes);  // ... those that exist
    const uint32_t bytes = (uint32_t)(p.s.bh * p.s.bw * sizeof(T));
    auto issue = [&](int pl) {
        const int k = pl - ra;
        mbar_expect_tx(&bar[k & 1], bytes);
        tma_load_4d(stage0 + (k & 1) * box_elems, &map, &bar[k & 1], x0 - p.s.c2, y0 - p.s.c1, pl + p.lo_planes, b);
    };
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    if (tid < PXB_ST2_MAXTAP + 2 * (C::R - 1)) {
        const int q = tid - (C::R - 1);
        c1s[tid] = (q >= 0 && q < p.s.k1) ? T(p.s.coef1[q]) : T(0);
    }
    __syncthreads();
    if (tid == 0) {
        if (ra < rb) issue(ra);
        if (ra + 1 < rb) issue(ra + 1);
    }
    T c2[NV * VEC - VEC + 1], c0v[K0];
    for (int q = 0; q < NV * VEC - VEC + 1; ++q) c2[q] = T(p.s.coef2[q]);
    for (int k = 0; k < K0; ++k) c0v[k] = T(p.coef0[k]);
    T ring[K0][C::R][VEC];
    for (int k = 0; k < K0; ++k)
        for (int r = 0; r < C::R; ++r)
            for (int j = 0; j < VEC; ++j) ring[k][r][j] = T(0);
    const int xl = (tid & 31) * VEC, yl = (tid >> 5) * C::R;

    for (int base = pl_lo; base < pl_hi; base += K0) {
#pragma unroll
        for (int u = 0; u < K0; ++u) {  // plane base+u -> ring slot u (compile-time after unrolling)
            const int pl = base + u;
            if (pl < pl_hi) {
                const bool have = pl >= ra && pl < rb;
                const int q = pl - (K0 - 1 - p.c0);
                T addv[C::R][VEC];
                if (q >= m0) pxb_st3_load_add<T, VEC>(p, addv, b, q, y0, x0, yl, xl);  // in flight during the passes
                if (have) {
                    const int k = pl - ra;
                    const T* box = stage0 + (k & 1) * box_elems;
                    mbar_wait(&bar[k & 1], (uint32_t)(k >> 1) & 1u);
                    for (int it = tid; it < p.s.bh * C::TXL; it += C::NT) pxb_st3_row_item<T, VEC, NV>(p, box, mid, it >> 5, (it & 31) * VEC, c2);
                    __syncthreads();  // intermediate complete; this stage's box is free
                    if (tid == 0 && pl + 2 < rb) issue(pl + 2);
                    pxb_st3_col_item<T, VEC>(p, mid, yl, xl, c1s, ring[u]);
                } else {
                    for (int r = 0; r < C::R; ++r)
                        for (int j = 0; j < VEC; ++j) ring[u][r][j] = T(0);
                }
                if (q >= m0) pxb_st3_store<T, VEC, K0>(p, out, c0v, ring, u, addv, b, q, y0, x0, yl, xl);
                if (have && pl + 1 < rb) __syncthreads();  // `mid` is rewritten by the next plane's row pass
            }
        }
    }
}

template <class T, int VEC, int NV, int K0>
cudaError_t launch(const PxbSt3P& p, const CUtensorMap& map, T* out, cudaStream_t s) {
    using C = PxbSt3Cfg<T, VEC>;
    const size_t box_bytes = (size_t)((p.s.bh * p.s.bw + 31) / 32 * 32) * sizeof(T);
    const size_t smem = 2 * box_bytes + (size_t)p.s.bh * C::TX * sizeof(T);
    auto k = k_stencil3d<T, VEC, NV, K0>;
    if (smem > 48 * 1024) {
        cudaError_t e = pxb_smem_attr_once((const void*)k, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const unsigned grid = (unsigned)((int64_t)p.s.ntx * p.s.nty * p.nchunk * p.batch);
    k<<<grid, C::NT, smem, s>>>(p, map, out);
    return cudaGetLastError();
}

template <class T, int VEC, int NV>
bool launch_k0(const PxbSt3P& p, const CUtensorMap& map, T* out, cudaStream_t s, cudaError_t* err) {
    switch (p.k0) {
        case 3: *err = launch<T, VEC, NV, 3>(p, map, out, s); return true;
        case 5: *err = launch<T, VEC, NV, 5>(p, map, out, s); return true;
        case 7: *err = launch<T, VEC, NV, 7>(p, map, out, s); return true;
        case 9: *err = launch<T, VEC, NV, 9>(p, map, out, s); return true;
        default: return false;
    }
}

template <class T>
int run(PxbSt3P& p, const void* in, void* out, cudaStream_t s, cudaError_t* err) {
    constexpr int VEC = 16 / (int)sizeof(T);
    if (int why = pxb_st3_setup<T, VEC>(p)) return why;
    if (g_st3_path == 0 && pxb_st3_fast_try(sizeof(T) == 4 ? PXB_F32 : PXB_F64, p, in, out, s, err) == 0) return 0;
    const int64_t s0 = (int64_t)p.s.n1 * p.s.n2;
    const uint64_t dim[4] = {(uint64_t)p.s.n2, (uint64_t)p.s.n1, (uint64_t)(p.n0 + p.lo_planes + p.hi_planes), (uint64_t)p.batch};
    const uint64_t stride[4] = {1, (uint64_t)p.s.n2, (uint64_t)s0, (uint64_t)p.vol};
    const uint32_t box[4] = {(uint32_t)p.s.bw, (uint32_t)p.s.bh, 1, 1};
    alignas(64) CUtensorMap map;
    if (!pxb_tma_encode_cached<T>(4, (const T*)in - (int64_t)p.lo_planes * s0, dim, stride, box, &map)) return 10;
    bool ok = false;
    switch (pxb_st2_nv(p.s.k2, VEC)) {
        case 1: ok = launch_k0<T, VEC, 1>(p, map, (T*)out, s, err); break;
        case 2: ok = launch_k0<T, VEC, 2>(p, map, (T*)out, s, err); break;
        case 3: ok = launch_k0<T, VEC, 3>(p, map, (T*)out, s, err); break;
        case 4: ok = launch_k0<T, VEC, 4>(p, map, (T*)out, s, err); break;
        default:
            if constexpr (VEC == 2) {
                const int nv = pxb_st2_nv(p.s.k2, VEC);
                if (nv == 5) ok = launch_k0<T, VEC, 5>(p, map, (T*)out, s, err);
                else if (nv == 6) ok = launch_k0<T, VEC, 6>(p, map, (T*)out, s, err);
            }
            break;
    }
    return ok ? 0 : 1;
}

}  // namespace

extern "C" int pxb_set_stencil3d_path(int path) {
    if (path < 0 || path > 1) return pxb_fail(PXB_EINVAL, "pxb_set_stencil3d_path: 0 (auto) or 1 (generic kernel)");
    g_st3_path = path;
    return 0;
}

extern "C" int pxb_stencil3d_apply(const pxb_stencil3d* d, const void* in, void* out, void* stream) {
    const char* who = "pxb_stencil3d_apply";
    if (!d || !in || !out || in == out) return pxb_fail(PXB_EINVAL, "%s: null or aliased argument", who);
    if (d->dtype != PXB_F32 && d->dtype != PXB_F64) return pxb_fail(PXB_EINVAL, "%s: bad dtype %d", who, d->dtype);
    if (d->batch < 1 || d->shape[0] < 1 || d->shape[1] < 1 || d->shape[2] < 1) return pxb_fail(PXB_EINVAL, "%s: empty array", who);
    for (int a = 0; a < 3; ++a)
        if (d->ksize[a] < 1 || d->ksize[a] > 16 || d->center[a] < 0 || d->center[a] >= d->ksize[a]) return pxb_fail(PXB_EINVAL, "%s: bad kernel extent / center along axis %d", who, a);
    const int halo = d->slab.halo;
    if (halo > 0 && d->batch != 1) return pxb_fail(PXB_EINVAL, "%s: slabs require batch == 1", who);
    PxbSt3P p;
    p.s.n1 = (int)d->shape[1]; p.s.n2 = (int)d->shape[2];
    p.s.k1 = d->ksize[1]; p.s.k2 = d->ksize[2]; p.s.c1 = d->center[1]; p.s.c2 = d->center[2];
    for (int i = 0; i < PXB_ST2_MAXTAP; ++i) { p.s.coef1[i] = d->coef1[i]; p.s.coef2[i] = d->coef2[i]; p.coef0[i] = d->coef0[i]; }
    p.s.coef = nullptr; p.s.alpha = d->alpha; p.s.beta = d->beta; p.s.add = d->add; p.s.add_period = d->add_period;
    if (d->add && d->add_period > 0 && d->add_period >= d->batch * d->shape[0] * d->shape[1] * d->shape[2]) p.s.add_period = 0;
    p.s.pa = 1.0; p.s.pb = 0.0; p.s.epi = 0; p.s.e1 = p.s.e2 = nullptr; p.s.norms = nullptr; p.s.imgs_per_row = 1;
    p.n0 = (int)d->shape[0]; p.batch = d->batch;
    const int alloc = d->slab.plane_alloc > 0 ? d->slab.plane_alloc : p.n0 + 2 * halo;
    p.vol = (int64_t)alloc * d->shape[1] * d->shape[2];
    p.k0 = d->ksize[0]; p.c0 = d->center[0];
    p.lo_planes = d->slab.open_lo ? p.c0 : 0;
    p.hi_planes = d->slab.open_hi ? p.k0 - 1 - p.c0 : 0;
    if (p.lo_planes > halo || p.hi_planes > halo) return pxb_fail(PXB_EINVAL, "%s: the stencil reaches %d / %d planes across an open side but halo = %d", who, p.lo_planes, p.hi_planes, halo);
    if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(d->add)) & 15u)
        return pxb_fail(PXB_ENOSUP, "%s: arrays must be 16-byte aligned", who);
    cudaError_t err = cudaSuccess;
    const int why = d->dtype == PXB_F32 ? run<float>(p, in, out, (cudaStream_t)stream, &err) : run<double>(p, in, out, (cudaStream_t)stream, &err);
    if (why) return pxb_fail(PXB_ENOSUP, "%s: outside the single-pass kernel's envelope (reason %d)", who, why);
    pxb_count_launch();
    if (err != cudaSuccess) return pxb_fail(PXB_ECUDA, "%s: %s", who, cudaGetErrorString(err));
    return 0;
}
