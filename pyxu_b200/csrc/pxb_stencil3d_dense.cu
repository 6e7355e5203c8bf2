// pxb_stencil3d_dense.cu -- launcher of the dense K x K x K stencil (design: pxb_stencil3d_dense.cuh).
#include "pxb_launch.cuh"
#include "pxb_stencil3d_dense.cuh"

namespace {

template <class T, int K>
__global__ void __launch_bounds__(256, 2) k_stencil3d_dense(const __grid_constant__ PxbD3P<T, K> p, const T* __restrict__ in, T* __restrict__ out) {
    using C = PxbD3Cfg<T, K>;
    constexpr int VEC = C::VEC;
    __shared__ __align__(16) T box[2][C::BOX];
    const int tid = threadIdx.x;
    unsigned blk = blockIdx.x;
    const int tx = blk % (unsigned)p.ntx; blk /= (unsigned)p.ntx;
    const int ty = blk % (unsigned)p.nty; blk /= (unsigned)p.nty;
    const int ch = blk % (unsigned)p.nchunk;
    const int64_t b = blk / (unsigned)p.nchunk;
    const int x0 = tx * C::TX, y0 = ty * C::TY;
    const int m0 = ch * p.chunk, m1 = min(p.n0, m0 + p.chunk);
    const int pl_lo = m0 - p.c0, pl_hi = m1 + K - 1 - p.c0;                        // input planes this chunk's outputs take
    const int ra = max(pl_lo, -p.lo_planes), rb = min(pl_hi, p.n0 + p.hi_planes);  // ... those that exist
    const int64_t s0 = (int64_t)p.n1 * p.n2;
    const T* __restrict__ vol = in + b * p.vol;
    const int xl = (tid & 31) * VEC, yl = (tid >> 5) * C::R;

    T acc[K][C::R][VEC];
#pragma unroll
    for (int a = 0; a < K; ++a)
        for (int r = 0; r < C::R; ++r)
            for (int j = 0; j < VEC; ++j) acc[a][r][j] = T(0);
    T pre[C::NROW * C::NCOL];
    if (ra < rb) {
        pxb_d3_fetch<T, K>(p, vol + (int64_t)ra * s0, y0, x0, tid, pre);
        pxb_d3_stash<T, K>(pre, box[0], tid);
    }
    __syncthreads();
    // planes below `ra` do not exist and no output is complete before plane ra: the march starts there
    for (int pl = ra; pl < pl_hi; ++pl) {
        const bool have = pl < rb, more = pl + 1 < rb;
        const int k = pl - ra;
        if (more) pxb_d3_fetch<T, K>(p, vol + (int64_t)(pl + 1) * s0, y0, x0, tid, pre);  // in flight during the accumulation
        const int q = pl - (K - 1 - p.c0);
        T addv[C::R][VEC];
        if (q >= m0) pxb_d3_load_add<T, K>(p, addv, b, q, y0, x0, yl, xl);
        if (have) {
            const int a_lo = pl + p.c0 - m1 + 1, a_hi = pl + p.c0 - m0;  // slots whose output plane pl - a + c0 lies in [m0, m1)
            if (a_lo <= 0 && a_hi >= K - 1) pxb_d3_accum<T, K>(p.coef, box[k & 1], yl, xl, acc);
            else pxb_d3_accum_some<T, K>(p.coef, box[k & 1], yl, xl, acc, a_lo, a_hi);
        }
        if (q >= m0) pxb_d3_emit<T, K>(p, out, acc[K - 1], addv, b, q, y0, x0, yl, xl);
        pxb_d3_shift<T, K>(acc);
        if (more) {
            pxb_d3_stash<T, K>(pre, box[(k + 1) & 1], tid);  // the other buffer: last read before the previous barrier
            __syncthreads();
        }
    }
}

template <class T, int K>
int run_k(const pxb_stencil3d_dense* d, const T* in, T* out, cudaStream_t s, cudaError_t* err) {
    using C = PxbD3Cfg<T, K>;
    static_assert(sizeof(PxbD3P<T, K>) <= 4000, "the coefficients travel as a kernel parameter");
    PxbD3P<T, K> p;
    p.n0 = (int)d->shape[0]; p.n1 = (int)d->shape[1]; p.n2 = (int)d->shape[2];
    p.batch = d->batch;
    const int halo = d->slab.halo;
    const int alloc = d->slab.plane_alloc > 0 ? d->slab.plane_alloc : p.n0 + 2 * halo;
    p.vol = (int64_t)alloc * d->shape[1] * d->shape[2];
    p.c0 = d->center[0]; p.c1 = d->center[1]; p.c2 = d->center[2];
    // ghost planes the kernel itself reaches (the zero taps of the enclosing cube read nothing)
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
    if (int why = pxb_d3_setup<T, K>(p)) return why;
    const unsigned grid = (unsigned)((int64_t)p.ntx * p.nty * p.nchunk * p.batch);
    k_stencil3d_dense<T, K><<<grid, C::NT, 0, s>>>(p, in, out);
    *err = cudaGetLastError();
    return 0;
}

template <class T>
int run(const pxb_stencil3d_dense* d, const void* in, void* out, cudaStream_t s, cudaError_t* err) {
    switch (pxb_d3_cube(d->ksize)) {
        case 3: return run_k<T, 3>(d, (const T*)in, (T*)out, s, err);
        case 5: return run_k<T, 5>(d, (const T*)in, (T*)out, s, err);
        case 7: return run_k<T, 7>(d, (const T*)in, (T*)out, s, err);
        default: return 1;
    }
}

}  // namespace

extern "C" int pxb_stencil3d_dense_apply(const pxb_stencil3d_dense* d, const void* in, void* out, void* stream) {
    const char* who = "pxb_stencil3d_dense_apply";
    if (!d || !in || !out || in == out || !d->coef) return pxb_fail(PXB_EINVAL, "%s: null or aliased argument", who);
    if (d->dtype != PXB_F32 && d->dtype != PXB_F64) return pxb_fail(PXB_EINVAL, "%s: bad dtype %d", who, d->dtype);
    if (d->batch < 1 || d->shape[0] < 1 || d->shape[1] < 1 || d->shape[2] < 1) return pxb_fail(PXB_EINVAL, "%s: empty array", who);
    if (d->shape[0] > 0x7fffffffLL || d->shape[1] > 0x7fffffffLL || d->shape[2] > 0x7fffffffLL) return pxb_fail(PXB_ENOSUP, "%s: axis longer than 2^31 - 1", who);
    for (int a = 0; a < 3; ++a)
        if (d->ksize[a] < 1 || d->center[a] < 0 || d->center[a] >= d->ksize[a]) return pxb_fail(PXB_EINVAL, "%s: bad kernel extent / center along axis %d", who, a);
    const int halo = d->slab.halo;
    if (halo > 0 && d->batch != 1) return pxb_fail(PXB_EINVAL, "%s: slabs require batch == 1", who);
    if ((d->slab.open_lo && d->center[0] > halo) || (d->slab.open_hi && d->ksize[0] - 1 - d->center[0] > halo))
        return pxb_fail(PXB_EINVAL, "%s: the stencil reaches %d / %d planes across an open side but halo = %d", who, d->center[0], d->ksize[0] - 1 - d->center[0], halo);
    if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(d->add)) & 15u)
        return pxb_fail(PXB_ENOSUP, "%s: arrays must be 16-byte aligned", who);
    cudaError_t err = cudaSuccess;
    const int why = d->dtype == PXB_F32 ? run<float>(d, in, out, (cudaStream_t)stream, &err) : run<double>(d, in, out, (cudaStream_t)stream, &err);
    if (why) return pxb_fail(PXB_ENOSUP, "%s: outside the dense marching kernel's envelope (reason %d)", who, why);
    pxb_count_launch();
    if (err != cudaSuccess) return pxb_fail(PXB_ECUDA, "%s: %s", who, cudaGetErrorString(err));
    return 0;
}
