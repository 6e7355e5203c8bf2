// pxb_tv_iter.cu -- launcher of the single-kernel PD3O / CondatVu iteration (see pxb_tv_iter.cuh for the design).
//
// Roofline (fp32, 3-D, per voxel and iteration): read u, shift, z0, z1, z2 and write u, z0, z1, z2 = 36 B
// (40 B when x is materialised, 44 B when RelError[x] also re-reads the previous x).  The rim samples of a
// tile and the planes shared by consecutive chunks are re-read through L1/L2, not HBM.
#include <initializer_list>

#include "pxb_launch.cuh"
#include "pxb_tma_util.cuh"
#include "pxb_tv_iter.cuh"

namespace {

// occupancy target: 3 CTAs of 256 threads per SM (<= 80 registers), each with ~20 KB of ring
template <class T, int VEC, int TXL, int TY, int NDIR, int ALGO, bool NORMS, bool MODES>
__global__ void __launch_bounds__(TXL* TY, (TXL * TY >= 256 ? 3 : 6))
    k_tv_iter(const __grid_constant__ PxbTvP<T> q, const __grid_constant__ PxbIterGeom g, const __grid_constant__ PxbIterPtr<T> a) {
    using C = PxbIterCfg<T, VEC, TXL, TY, NDIR>;
    if (NORMS && pxb_iter_stopped(a.stop)) return;  // an earlier iteration of this batch met the stopping rule
    extern __shared__ __align__(16) unsigned char pxb_iter_smem[];
    T* smem = reinterpret_cast<T*>(pxb_iter_smem);
    const PxbIterItem it = pxb_iter_item(g, (int64_t)blockIdx.x, TY, C::T2);
    const PxbIterRange R = pxb_iter_range<T>(q, it);
    const int tid = threadIdx.x;
    PxbIterThread<T, VEC> st;
    for (int k = 0; k < 3; ++k)
        for (int j = 0; j < VEC; ++j) st.zc[k][j] = st.zprev[k][j] = T(0);
    for (int k = 0; k < 4; ++k) st.acc[k] = 0.0;

    for (int m = R.mlo; m < R.mhi; ++m) {
        pxb_iter_phaseA<T, VEC, TXL, TY, NDIR, ALGO, NORMS, MODES>(q, g, it, a, tid, m, smem, st);
        __syncthreads();
        const int mm = m - R.lag;
        if (mm >= it.m0 && mm < it.m1) {
            T zo[3][VEC];
            for (int k = 0; k < 3; ++k)
                for (int j = 0; j < VEC; ++j) zo[k][j] = R.lag ? st.zprev[k][j] : st.zc[k][j];
            pxb_iter_phaseC<T, VEC, TXL, TY, NDIR, NORMS>(q, g, it, a, tid, mm, smem, zo, st.acc);
        }
        for (int k = 0; k < 3; ++k)
            for (int j = 0; j < VEC; ++j) st.zprev[k][j] = st.zc[k][j];
    }

    if (NORMS) {  // every thread of the CTA works on the same batch row: warp shuffle -> shared -> one atomic per sum
        __shared__ double red[4][C::NT / 32];
        double v[4] = {st.acc[0], st.acc[1], st.acc[2], st.acc[3]};
        for (int o = 16; o > 0; o >>= 1)
            for (int k = 0; k < 4; ++k) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
        const int w = tid >> 5, l = tid & 31;
        if (l == 0)
            for (int k = 0; k < 4; ++k) red[k][w] = v[k];
        __syncthreads();
        if (w == 0) {
            for (int k = 0; k < 4; ++k) {
                double s = 0.0;
                for (int i = l; i < C::NT / 32; i += 32) s += red[k][i];
                for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
                v[k] = s;
            }
            if (l == 0) {
                if (a.norms_x) { atomicAdd(a.norms_x + 2 * it.b, v[0]); atomicAdd(a.norms_x + 2 * it.b + 1, v[1]); }
                if (a.norms_z) { atomicAdd(a.norms_z + 2 * it.b, v[2]); atomicAdd(a.norms_z + 2 * it.b + 1, v[3]); }
                pxb_iter_finish(a.stop, a.norms_x, a.norms_z, gridDim.x);
            }
        }
    }
}

template <class T, int VEC, int TXL, int TY, int NDIR, int ALGO, bool NORMS, bool MODES>
cudaError_t launch_inst(const PxbTvP<T>& q, const PxbIterGeom& g, const PxbIterPtr<T>& a, cudaStream_t s) {
    using C = PxbIterCfg<T, VEC, TXL, TY, NDIR>;
    auto kern = k_tv_iter<T, VEC, TXL, TY, NDIR, ALGO, NORMS, MODES>;
    if (C::SMEM > 48 * 1024) {
        cudaError_t e = pxb_smem_attr_once((const void*)kern, (int)C::SMEM);
        if (e != cudaSuccess) return e;
    }
    kern<<<(unsigned)g.nblocks, C::NT, C::SMEM, s>>>(q, g, a);
    return cudaGetLastError();
}
// folding boundary modes run the MODES instance (fold terms of K^T z, folded rim of the w tiles: pxb_tv_iter.cuh)
template <class T, int VEC, int TXL, int TY, int NDIR, int ALGO, bool NORMS>
cudaError_t launch_cfg(const pxb_grad_desc& d, const PxbTvP<T>& q, const PxbIterGeom& g, const PxbIterPtr<T>& a, cudaStream_t s) {
    if (pxb_any_mode(d)) return launch_inst<T, VEC, TXL, TY, NDIR, ALGO, NORMS, true>(q, g, a, s);
    return launch_inst<T, VEC, TXL, TY, NDIR, ALGO, NORMS, false>(q, g, a, s);
}

template <class T>
struct Tile {  // 16-byte vectors; 3-D: one warp per row, 8 rows; 2-D: the whole CTA along the row
    static constexpr int VEC = 16 / (int)sizeof(T);
};

template <class T, int NDIR, int ALGO, bool NORMS>
int run(const pxb_grad_desc& d, const pxb_pds_params& P, const PxbIterPtr<T>& a, int chunk_hint, cudaStream_t s, cudaError_t* err) {
    constexpr int VEC = Tile<T>::VEC;
    PxbTvCoef cf;
    PxbIterGeom g;
    PxbTvP<T> q;
    if (NDIR == 3) {
        constexpr int TXL = 32, TY = 8;
        if (int why = pxb_iter_setup(d, P, VEC, TY, TXL * VEC, chunk_hint, 148 * 3, cf, g, pxb_iter_modes() != 0)) return why;
        pxb_tv_prepare<T>(d, cf, P, q);
        *err = launch_cfg<T, VEC, TXL, TY, 3, ALGO, NORMS>(d, q, g, a, s);
    } else {
        const bool narrow = d.shape[2] <= 128 * VEC;
        if (int why = pxb_iter_setup(d, P, VEC, 1, (narrow ? 128 : 256) * VEC, chunk_hint, 148 * 3, cf, g, pxb_iter_modes() != 0)) return why;
        pxb_tv_prepare<T>(d, cf, P, q);
        *err = narrow ? launch_cfg<T, VEC, 128, 1, 2, ALGO, NORMS>(d, q, g, a, s) : launch_cfg<T, VEC, 256, 1, 2, ALGO, NORMS>(d, q, g, a, s);
    }
    return 0;
}

template <class T>
int dispatch(int algo, const pxb_grad_desc& d, const pxb_pds_params& P, const PxbIterPtr<T>& a, int chunk_hint, cudaStream_t s,
             cudaError_t* err) {
    const bool norms = a.norms_x || a.norms_z;
#define PXB_ITER_CASE(ND, AL)                                                                       \
    if (d.ndir == ND && algo == AL)                                                                 \
        return norms ? run<T, ND, AL, true>(d, P, a, chunk_hint, s, err) : run<T, ND, AL, false>(d, P, a, chunk_hint, s, err);
    PXB_ITER_CASE(3, PXB_PD3O)
    PXB_ITER_CASE(3, PXB_CV)
    PXB_ITER_CASE(2, PXB_PD3O)
    PXB_ITER_CASE(2, PXB_CV)
#undef PXB_ITER_CASE
    return 2;
}

bool aligned16(std::initializer_list<const void*> ptrs) {
    for (const void* p : ptrs)
        if (p && (reinterpret_cast<uintptr_t>(p) & 15u)) return false;
    return true;
}

}  // namespace

// 0: launched; PXB_ENOSUP: the descriptor is outside the envelope of the single-kernel form (the caller falls
// back to pxb_pds_primal + pxb_pds_dual); other negative codes: errors.
int pxb_tv_iter_launch(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, const void* xu_in, const void* z_in, void* xu_out,
                       void* z_out, void* x_out, double* norms_x, double* norms_z, int chunk_hint, cudaStream_t s, const PxbIterStop* stop,
                       const pxb_peer* peer) {
    const void* sh = (p->f.kind == PXB_F_SQL2 && p->f.shift_period > 1) ? p->f.shift : nullptr;
    const void* ga = p->f.kind == PXB_F_GRADARR ? p->f.garr : nullptr;
    if (!aligned16({xu_in, z_in, xu_out, z_out, x_out, sh, ga})) return pxb_fail(PXB_ENOSUP, "pxb_pds_iter: arrays must be 16-byte aligned");
    if (algo == PXB_PD3O && p->f.kind == PXB_F_GRADARR) return pxb_fail(PXB_ENOSUP, "pxb_pds_iter: PD3O needs grad f at the new x (pointwise f only)");
    if (algo == PXB_PD3O && norms_x && !x_out) return pxb_fail(PXB_EINVAL, "pxb_pds_iter: RelError[x] needs x_out (it holds the previous x)");
    cudaError_t err = cudaSuccess;
    int why;
    // 3-D volumes: TMA-staged pipeline (pxb_tv_tma.cu) unless the direct-load form is forced or the TMA form declines
    if (peer && (K->ndir != 3 || pxb_iter_path() == 1)) return pxb_fail(PXB_ENOSUP, "pxb_pds_iter_p2p: only the TMA form of 3-D volumes carries the peer-memory exchange");
    if (K->ndir == 3 && pxb_iter_path() != 1) {
        why = pxb_tv_tma_try(algo, K, p, xu_in, z_in, xu_out, z_out, x_out, norms_x, norms_z, chunk_hint, s, &err, stop, peer);
        if (why == 0) {
            pxb_count_launch();
            if (err != cudaSuccess) return pxb_fail(PXB_ECUDA, "pxb_pds_iter (tma): %s", cudaGetErrorString(err));
            return 0;
        }
        if (pxb_iter_path() == 2 || peer) return pxb_fail(PXB_ENOSUP, "pxb_pds_iter: TMA form not applicable (reason %d)", why);
    } else if (K->ndir == 2 && pxb_iter_path() != 1 && chunk_hint == 0) {
        // 2-D images: TMA-staged tiles (pxb_tv_tile2d.cu); the marching direct-load form remains the fallback
        why = pxb_tv_tile2d_try(algo, K, p, xu_in, z_in, xu_out, z_out, x_out, norms_x, norms_z, s, &err, stop);
        if (why == 0) {
            pxb_count_launch();
            if (err != cudaSuccess) return pxb_fail(PXB_ECUDA, "pxb_pds_iter (tile2d): %s", cudaGetErrorString(err));
            return 0;
        }
        if (pxb_iter_path() == 2) return pxb_fail(PXB_ENOSUP, "pxb_pds_iter: TMA form not applicable (reason %d)", why);
    } else if (pxb_iter_path() == 2) {
        return pxb_fail(PXB_ENOSUP, "pxb_pds_iter: TMA form not applicable");
    }
    const PxbIterStop st = stop ? *stop : PxbIterStop{};
    if (K->dtype == PXB_F32) {
        PxbIterPtr<float> a{(const float*)xu_in, (const float*)z_in, (float*)xu_out, (float*)z_out, (float*)x_out, norms_x, norms_z, st};
        why = dispatch<float>(algo, *K, *p, a, chunk_hint, s, &err);
    } else {
        PxbIterPtr<double> a{(const double*)xu_in, (const double*)z_in, (double*)xu_out, (double*)z_out, (double*)x_out, norms_x, norms_z, st};
        why = dispatch<double>(algo, *K, *p, a, chunk_hint, s, &err);
    }
    if (why) return pxb_fail(PXB_ENOSUP, "pxb_pds_iter: not eligible for the single-kernel iteration (reason %d)", why);
    pxb_count_launch();
    if (err != cudaSuccess) return pxb_fail(PXB_ECUDA, "pxb_pds_iter: %s", cudaGetErrorString(err));
    return 0;
}

// n iterations queued back to back on the stream, alternating the two (xu, z) pairs; see pxb_pds_iter_n in the header.
int pxb_tv_iter_launch_n(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, void* xu_a, void* z_a, void* xu_b, void* z_b, void* x_out,
                         double* norms, int n, const pxb_stop_rule* rule, void* ctl, cudaStream_t s) {
    PxbIterStop st{};
    if (rule) {  // (no rule: n plain iterations -- a criterion the host alone evaluates, e.g. MaxIter)
        st.ctl = (int32_t*)ctl;
        st.eps_x = rule->eps_x; st.eps_z = rule->eps_z;
        st.all_x = rule->all_x; st.all_z = rule->all_z;
        st.table = rule->table;
        st.rows = (int32_t)K->batch;
    }
    const int64_t per = 4 * K->batch;
    // small 2-D problems (every tile resident at once): all n iterations in ONE cooperative launch with grid-wide barriers
    if (K->ndir == 2 && n >= 2 && pxb_iter_path() != 1) {
        const void* sh = (p->f.kind == PXB_F_SQL2 && p->f.shift_period > 1) ? p->f.shift : nullptr;
        const bool ok = aligned16({xu_a, z_a, xu_b, z_b, x_out, sh, p->f.kind == PXB_F_GRADARR ? p->f.garr : nullptr}) &&
                        !(algo == PXB_PD3O && p->f.kind == PXB_F_GRADARR) && !(algo == PXB_PD3O && rule && rule->eps_x > 0 && !x_out);
        if (ok) {
            cudaError_t err = cudaSuccess;
            const int why = pxb_tv_tile2d_loop_try(algo, K, p, xu_a, z_a, xu_b, z_b, x_out, rule ? norms : nullptr, rule && rule->eps_x > 0, rule && rule->eps_z > 0,
                                                   n, s, &err, rule ? &st : nullptr);
            if (why == 0) {
                if (err != cudaSuccess) return pxb_fail(PXB_ECUDA, "pxb_pds_iter_n (persistent tile2d): %s", cudaGetErrorString(err));
                pxb_count_launch();  // ONE kernel launch carries the n iterations
                return 0;
            }
        }
    }
    for (int i = 0; i < n; ++i) {
        double* nx = rule && rule->eps_x > 0 ? norms + (int64_t)i * per : nullptr;
        double* nz = rule && rule->eps_z > 0 ? norms + (int64_t)i * per + 2 * K->batch : nullptr;
        const bool even = (i & 1) == 0;
        int rc = pxb_tv_iter_launch(algo, K, p, even ? xu_a : xu_b, even ? z_a : z_b, even ? xu_b : xu_a, even ? z_b : z_a, x_out, nx, nz, 0, s, rule ? &st : nullptr);
        if (rc != 0) return i == 0 ? rc : pxb_fail(PXB_ECUDA, "pxb_pds_iter_n: launch %d of %d failed after earlier ones succeeded", i, n);
    }
    return 0;
}
