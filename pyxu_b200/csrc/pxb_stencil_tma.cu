// pxb_stencil_tma.cu -- launcher of the TMA-staged 2-D stencil (design: pxb_stencil_tma.cuh).
//   thread 0:   cp.async.bulk.tensor.3d box loads (SASS UTMALDG) of the input windows, mbarrier completion; a CTA walks up
//               to 4 consecutive tiles of a tile row with the next tile's box in flight (two stages)
//   256 threads: [dense] stage the k1 x k2 coefficients in shared memory while the first box is in flight;
//                wait -> row pass -> __syncthreads -> column pass + epilogue   (separable)
//                wait -> register-blocked accumulation + epilogue              (dense)
#include "pxb_launch.cuh"
#include "pxb_tma_util.cuh"
#include "pxb_stencil_tma.cuh"

namespace {

// MODE 0: out = alpha*S(in) + beta*add      1: the same on the window pa*in + pb*in2 (two box loads, combined in
// shared memory)      2: proximal-gradient epilogue (pxb_st2_store_prox)
// One tile per CTA, no in-CTA pipelining: the dense (FMA-bound) instances.  Walking several tiles per CTA (below) costs them
// occupancy and wave balance: measured 8192^2 dense 9x9 0.264 ms here vs 0.28-0.31 ms through the pipelined kernel.
template <class T, int VEC, int NV, bool DENSE, int MODE>
__global__ void __launch_bounds__(256) k_stencil2d_one(const __grid_constant__ PxbSt2P p, const __grid_constant__ CUtensorMap map,
                                                       const __grid_constant__ CUtensorMap map2, T* __restrict__ out) {
    using C = PxbSt2Cfg<T, VEC>;
    extern __shared__ __align__(128) unsigned char pxb_st2_smem[];
    __shared__ __align__(8) uint64_t bar;
    T* box = reinterpret_cast<T*>(pxb_st2_smem);
    const int box_elems = (p.bh * p.bw + 31) / 32 * 32;
    T* box2 = box + box_elems;                       // MODE 1 only
    T* mid = box + (MODE == 1 ? 2 : 1) * box_elems;  // separable: bh x TX intermediate; dense: k1*k2 coefficients
    const int tid = threadIdx.x;
    unsigned blk = blockIdx.x;
    const int tx = blk % (unsigned)p.ntx; blk /= (unsigned)p.ntx;
    const int ty = blk % (unsigned)p.nty;
    const int64_t img = blk / (unsigned)p.nty;
    const int x0 = tx * C::TX, y0 = ty * C::TY;
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)(p.bh * p.bw * sizeof(T));
        mbar_expect_tx(&bar, MODE == 1 ? 2 * bytes : bytes);
        tma_load_3d(box, &map, &bar, x0 - p.c2, y0 - p.c1, (int)img);
        if (MODE == 1) tma_load_3d(box2, &map2, &bar, x0 - p.c2, y0 - p.c1, (int)img);
    }
    if (DENSE) {
        const T* __restrict__ ck = (const T*)p.coef;
        for (int i = tid; i < p.k1 * p.k2; i += C::NT) mid[i] = pxb_st2_dense_coef<T>(p, ck, i);
    }
    const int xl = (tid & 31) * VEC, yl = (tid >> 5) * C::R;
    PxbSt2Epi<T, VEC> epi;
    if (DENSE) pxb_st2_prefetch_epi<T, VEC>(p, img, y0, x0, yl, xl);   // towards L2 while the box arrives
    else pxb_st2_load_epi<T, VEC>(p, epi, img, y0, x0, yl, xl);        // into registers, in flight while the box arrives
    mbar_wait(&bar, 0);
    if (MODE == 1) {
        for (int it = tid; it < p.bh * p.bw / VEC; it += C::NT) pxb_st2_combine_item<T, VEC>(p, box, box2, it);
        __syncthreads();
    }
    T acc[C::R][VEC];
    if (!DENSE) {
        __shared__ T c1s[PXB_ST2_MAXTAP + 2 * (C::R - 1)];  // row factor with R-1 zeros on each side
        if (tid < PXB_ST2_MAXTAP + 2 * (C::R - 1)) {
            const int q = tid - (C::R - 1);
            c1s[tid] = (q >= 0 && q < p.k1) ? T(p.coef1[q]) : T(0);
        }
        T c2[NV * VEC - VEC + 1];
        for (int q = 0; q < NV * VEC - VEC + 1; ++q) c2[q] = T(p.coef2[q]);
        for (int it = tid; it < p.bh * C::TXL; it += C::NT) pxb_st2_row_item<T, VEC, NV>(p, box, mid, it >> 5, (it & 31) * VEC, c2);
        __syncthreads();
        pxb_st2_col_item<T, VEC>(p, mid, yl, xl, c1s, acc);
    } else {
        __syncthreads();
        pxb_st2_dense_item<T, VEC, NV>(p, box, mid, yl, xl, acc);
        pxb_st2_load_epi<T, VEC>(p, epi, img, y0, x0, yl, xl);
    }
    if (MODE != 2) {
        pxb_st2_store<T, VEC>(p, out, epi, img, y0, x0, yl, xl, acc);
    } else {
        double nrm[2] = {0.0, 0.0};
        pxb_st2_store_prox<T, VEC>(p, out, epi, img, y0, x0, yl, xl, acc, nrm);
        if (p.norms) {  // one image per CTA: warp shuffle -> shared -> one atomic pair
            __shared__ double red[2][C::NT / 32];
            for (int o = 16; o > 0; o >>= 1) {
                nrm[0] += __shfl_down_sync(0xffffffffu, nrm[0], o);
                nrm[1] += __shfl_down_sync(0xffffffffu, nrm[1], o);
            }
            if ((tid & 31) == 0) { red[0][tid >> 5] = nrm[0]; red[1][tid >> 5] = nrm[1]; }
            __syncthreads();
            if (tid == 0) {
                double s0 = 0.0, s1 = 0.0;
                for (int i = 0; i < C::NT / 32; ++i) { s0 += red[0][i]; s1 += red[1][i]; }
                const int64_t row = img / p.imgs_per_row;
                atomicAdd(p.norms + 2 * row, s0);
                atomicAdd(p.norms + 2 * row + 1, s1);
            }
        }
    }
}

// Several consecutive tiles per CTA with the next tile's box in flight: the separable (load-latency-bound) instances.
template <class T, int VEC, int NV, bool DENSE, int MODE>
__global__ void __launch_bounds__(256, 4) k_stencil2d_tma(const __grid_constant__ PxbSt2P p, const __grid_constant__ CUtensorMap map,
                                                       const __grid_constant__ CUtensorMap map2, T* __restrict__ out) {
    using C = PxbSt2Cfg<T, VEC>;
    constexpr int NBOX = MODE == 1 ? 2 : 1;  // boxes per stage
    extern __shared__ __align__(128) unsigned char pxb_st2_smem[];
    __shared__ __align__(8) uint64_t bar[2];
    T* stage0 = reinterpret_cast<T*>(pxb_st2_smem);
    const int box_elems = (p.bh * p.bw + 31) / 32 * 32;
    T* mid = stage0 + (p.tpc > 1 ? 2 : 1) * NBOX * box_elems;  // separable: bh x TX intermediate; dense: k1*k2 coefficients
    const int tid = threadIdx.x;
    unsigned blk = blockIdx.x;
    const int gx = blk % (unsigned)p.ngx; blk /= (unsigned)p.ngx;
    const int ty = blk % (unsigned)p.nty;
    const int64_t img = blk / (unsigned)p.nty;
    const int tx0 = gx * p.tpc, ntile = min(p.tpc, p.ntx - tx0);
    const int y0 = ty * C::TY;
    const uint32_t bytes = (uint32_t)(p.bh * p.bw * sizeof(T));

    // a CTA walks `ntile` consecutive tiles of one tile row; the box of tile t+1 is in flight while tile t is computed
    auto issue = [&](int t) {
        T* st = stage0 + (t & 1) * NBOX * box_elems;
        mbar_expect_tx(&bar[t & 1], NBOX * bytes);
        const int xs = (tx0 + t) * C::TX - p.c2;
        tma_load_3d(st, &map, &bar[t & 1], xs, y0 - p.c1, (int)img);
        if (MODE == 1) tma_load_3d(st + box_elems, &map2, &bar[t & 1], xs, y0 - p.c1, (int)img);
    };
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        issue(0);
        if (ntile > 1) issue(1);
    }
    if (DENSE) {
        const T* __restrict__ ck = (const T*)p.coef;
        for (int i = tid; i < p.k1 * p.k2; i += C::NT) mid[i] = pxb_st2_dense_coef<T>(p, ck, i);
    }
    __shared__ T c1s[PXB_ST2_MAXTAP + 2 * (C::R - 1)];  // row factor with R-1 zeros on each side
    if (!DENSE && tid < PXB_ST2_MAXTAP + 2 * (C::R - 1)) {
        const int q = tid - (C::R - 1);
        c1s[tid] = (q >= 0 && q < p.k1) ? T(p.coef1[q]) : T(0);
    }
    T c2[NV * VEC - VEC + 1];
    if (!DENSE)
        for (int q = 0; q < NV * VEC - VEC + 1; ++q) c2[q] = T(p.coef2[q]);
    const int xl = (tid & 31) * VEC, yl = (tid >> 5) * C::R;
    double nrm[2] = {0.0, 0.0};

    for (int t = 0; t < ntile; ++t) {
        T* box = stage0 + (t & 1) * NBOX * box_elems;
        const int x0 = (tx0 + t) * C::TX;
        PxbSt2Epi<T, VEC> epi;
        pxb_st2_load_epi<T, VEC>(p, epi, img, y0, x0, yl, xl);  // in flight while the box arrives / the passes run
        mbar_wait(&bar[t & 1], (uint32_t)(t >> 1) & 1u);
        if (MODE == 1) {
            for (int it = tid; it < p.bh * p.bw / VEC; it += C::NT) pxb_st2_combine_item<T, VEC>(p, box, box + box_elems, it);
            __syncthreads();
        }
        T acc[C::R][VEC];
        if (!DENSE) {
            for (int it = tid; it < p.bh * C::TXL; it += C::NT) pxb_st2_row_item<T, VEC, NV>(p, box, mid, it >> 5, (it & 31) * VEC, c2);
            __syncthreads();                                  // intermediate complete; nobody reads this stage's box any more
            if (tid == 0 && t + 2 < ntile) issue(t + 2);
            pxb_st2_col_item<T, VEC>(p, mid, yl, xl, c1s, acc);
        } else {
            if (t == 0) __syncthreads();                      // coefficients staged
            pxb_st2_dense_item<T, VEC, NV>(p, box, mid, yl, xl, acc);
        }
        if (MODE != 2) pxb_st2_store<T, VEC>(p, out, epi, img, y0, x0, yl, xl, acc);
        else pxb_st2_store_prox<T, VEC>(p, out, epi, img, y0, x0, yl, xl, acc, nrm);
        if (t + 1 < ntile) {
            __syncthreads();                                  // separable: `mid` is free again; dense: this stage's box is
            if (DENSE && tid == 0 && t + 2 < ntile) issue(t + 2);
        }
    }
    if (MODE == 2 && p.norms) {  // one image per CTA: warp shuffle -> shared -> one atomic pair
        __shared__ double red[2][C::NT / 32];
        for (int o = 16; o > 0; o >>= 1) {
            nrm[0] += __shfl_down_sync(0xffffffffu, nrm[0], o);
            nrm[1] += __shfl_down_sync(0xffffffffu, nrm[1], o);
        }
        if ((tid & 31) == 0) { red[0][tid >> 5] = nrm[0]; red[1][tid >> 5] = nrm[1]; }
        __syncthreads();
        if (tid == 0) {
            double s0 = 0.0, s1 = 0.0;
            for (int i = 0; i < C::NT / 32; ++i) { s0 += red[0][i]; s1 += red[1][i]; }
            const int64_t row = img / p.imgs_per_row;
            atomicAdd(p.norms + 2 * row, s0);
            atomicAdd(p.norms + 2 * row + 1, s1);
        }
    }
}

template <class T, int VEC, int NV, bool DENSE, int MODE>
cudaError_t launch_mode(const PxbSt2P& p, const CUtensorMap& map, const CUtensorMap& map2, T* out, cudaStream_t s) {
    using C = PxbSt2Cfg<T, VEC>;
    const size_t box_bytes = (size_t)((p.bh * p.bw + 31) / 32 * 32) * sizeof(T);
    const size_t smem = (p.tpc > 1 ? 2 : 1) * (MODE == 1 ? 2 : 1) * box_bytes + (DENSE ? (size_t)p.k1 * p.k2 * sizeof(T) : (size_t)p.bh * C::TX * sizeof(T));
    const unsigned grid = (unsigned)((int64_t)p.ngx * p.nty * p.nimg);
    auto k = p.tpc > 1 ? k_stencil2d_tma<T, VEC, NV, DENSE, MODE> : k_stencil2d_one<T, VEC, NV, DENSE, MODE>;
    if (smem > 48 * 1024) {
        cudaError_t e = pxb_smem_attr_once((const void*)k, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k<<<grid, C::NT, smem, s>>>(p, map, map2, out);
    return cudaGetLastError();
}

template <class T, int VEC, int NV>
cudaError_t launch_nv(const PxbSt2P& p, const CUtensorMap& map, const CUtensorMap& map2, int mode, T* out, cudaStream_t s) {
    if (p.dense) {
        if (mode == 1) return launch_mode<T, VEC, NV, true, 1>(p, map, map2, out, s);
        if (mode == 2) return launch_mode<T, VEC, NV, true, 2>(p, map, map2, out, s);
        return launch_mode<T, VEC, NV, true, 0>(p, map, map2, out, s);
    }
    if (mode == 1) return launch_mode<T, VEC, NV, false, 1>(p, map, map2, out, s);
    if (mode == 2) return launch_mode<T, VEC, NV, false, 2>(p, map, map2, out, s);
    return launch_mode<T, VEC, NV, false, 0>(p, map, map2, out, s);
}

template <class T>
int run(PxbSt2P& p, const PxbSt2In* ext, const void* in, const void* in2, void* out, cudaStream_t s, cudaError_t* err) {
    constexpr int VEC = 16 / (int)sizeof(T);
    if (int why = pxb_st2_setup<T, VEC>(p, ext)) return why;
    const uint64_t in_n1 = ext ? ext->n1 : p.n1, in_n2 = ext ? ext->n2 : p.n2;
    const uint64_t dim[3] = {in_n2, in_n1, (uint64_t)p.nimg};
    const uint64_t stride[3] = {1, in_n2, in_n1 * in_n2};
    const uint32_t box[3] = {(uint32_t)p.bw, (uint32_t)p.bh, 1};
    alignas(64) CUtensorMap map, map2;
    if (!pxb_tma_encode_cached<T>(3, in, dim, stride, box, &map)) return 10;
    if (!pxb_tma_encode_cached<T>(3, in2 ? in2 : in, dim, stride, box, &map2)) return 10;
    const int mode = in2 ? 1 : (p.epi == 1 ? 2 : 0);
    const int nv = pxb_st2_nv(p.k2, VEC);
    switch (nv) {
        case 1: *err = launch_nv<T, VEC, 1>(p, map, map2, mode, (T*)out, s); break;
        case 2: *err = launch_nv<T, VEC, 2>(p, map, map2, mode, (T*)out, s); break;
        case 3: *err = launch_nv<T, VEC, 3>(p, map, map2, mode, (T*)out, s); break;
        case 4: *err = launch_nv<T, VEC, 4>(p, map, map2, mode, (T*)out, s); break;
        default:
            if constexpr (VEC == 2) {
                if (nv == 5) { *err = launch_nv<T, VEC, 5>(p, map, map2, mode, (T*)out, s); break; }
                if (nv == 6) { *err = launch_nv<T, VEC, 6>(p, map, map2, mode, (T*)out, s); break; }
            }
            return 1;
    }
    return 0;
}

// fills the kernel parameter block from the C-ABI descriptors (shared with tests/emu through the header? no: host only);
// true when the input has an extent of its own (`ext`)
bool fill(PxbSt2P& p, PxbSt2In& ext, const pxb_stencil2d* d) {
    p.n1 = (int)d->shape[0]; p.n2 = (int)d->shape[1]; p.nimg = d->nimg;
    p.k1 = d->ksize[0]; p.k2 = d->ksize[1]; p.c1 = d->center[0]; p.c2 = d->center[1];
    const bool own_extent = d->in_shape[0] > 0 && d->in_shape[1] > 0;
    if (own_extent) {  // fold the origin of the output grid into the centers: the kernels only use them to place their windows
        ext.n1 = (int)d->in_shape[0]; ext.n2 = (int)d->in_shape[1];
        p.c1 -= d->origin[0]; p.c2 -= d->origin[1];
    }
    p.dense = d->dense;
    for (int i = 0; i < PXB_ST2_MAXTAP; ++i) { p.coef1[i] = d->coef1[i]; p.coef2[i] = d->coef2[i]; }
    p.coef = d->coef; p.alpha = d->alpha; p.beta = d->beta; p.add = d->add; p.add_period = d->add_period;
    if (d->add && d->add_period > 0 && d->add_period >= d->nimg * d->shape[0] * d->shape[1]) p.add_period = 0;
    p.pa = 1.0; p.pb = 0.0; p.epi = 0; p.e1 = p.e2 = nullptr; p.ea = p.eb = 0.0; p.gkind = 0; p.gp0 = p.gp1 = p.tau = 0.0;
    p.norms = nullptr; p.imgs_per_row = 1;
    return own_extent;
}

}  // namespace

// > 0: outside the envelope (reason); 0: launched or *err set
int pxb_stencil2d_try(const pxb_stencil2d* d, const void* in, void* out, cudaStream_t s, cudaError_t* err) {
    if (d->center[0] < 0 || d->center[0] >= d->ksize[0] || d->center[1] < 0 || d->center[1] >= d->ksize[1]) return 3;
    PxbSt2P p;
    PxbSt2In ext;
    const PxbSt2In* pe = fill(p, ext, d) ? &ext : nullptr;
    if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(d->add)) & 15u) return 6;
    return d->dtype == PXB_F32 ? run<float>(p, pe, in, nullptr, out, s, err) : run<double>(p, pe, in, nullptr, out, s, err);
}

int pxb_stencil2d_fista_try(const pxb_stencil2d* d, const pxb_fista_step* f, int which, void* out, cudaStream_t s, cudaError_t* err) {
    if (d->in_shape[0] > 0 || d->in_shape[1] > 0) return 7;  // the fused proximal-gradient passes keep input and output on one grid
    PxbSt2P p;
    PxbSt2In ext;
    fill(p, ext, d);
    if ((reinterpret_cast<uintptr_t>(f->x) | reinterpret_cast<uintptr_t>(f->x_prev) | reinterpret_cast<uintptr_t>(f->r) | reinterpret_cast<uintptr_t>(out) |
         reinterpret_cast<uintptr_t>(d->add)) & 15u)
        return 6;
    const void *in, *in2 = nullptr;
    if (which == 0) {  // r = alpha * S((1+a) x - a x_prev) + beta * add
        in = f->x;
        if (f->a != 0.0) { in2 = f->x_prev; p.pa = 1.0 + f->a; p.pb = -f->a; }
    } else {           // x_new = prox_{tau g}((1+a) x - a x_prev + alpha * S(r))
        in = f->r;
        p.epi = 1;
        p.e1 = f->x; p.e2 = f->a != 0.0 ? f->x_prev : nullptr;
        p.ea = 1.0 + f->a; p.eb = -f->a;
        p.gkind = f->g.kind; p.gp0 = f->g.p0; p.gp1 = f->g.p1; p.tau = f->tau;
        p.norms = f->norms; p.imgs_per_row = f->imgs_per_row > 0 ? f->imgs_per_row : 1;
    }
    return d->dtype == PXB_F32 ? run<float>(p, nullptr, in, in2, out, s, err) : run<double>(p, nullptr, in, in2, out, s, err);
}
