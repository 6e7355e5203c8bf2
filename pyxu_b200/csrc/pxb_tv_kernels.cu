// pxb_tv_kernels.cu -- the two kernels that carry one PD3O / CondatVu iteration of a TV-type problem
// (h o K with K a first-order finite-difference Gradient).  See pxb_tv_fast.cuh for the per-thread body.
//
// Roofline (fp32, 3-D, per voxel): primal reads u, y, z0, z1, z2 and writes x, w, u  -> 32 B;
//                                  dual   reads w, z0, z1, z2     and writes z0..z2 -> 28 B.
// Everything else (the -1/+1 neighbours) is served by L1/L2.  One thread = 16 bytes of every operand; the
// register budget is capped (launch bounds) so that >= 4 blocks of 256 threads stay resident per SM: with
// ~9 independent 128-bit loads per thread that keeps well over the ~35 KB/SM in flight that 6.5 TB/s needs.
#include <initializer_list>

#include "pxb_launch.cuh"
#include "pxb_tv_fast.cuh"

namespace {

constexpr int kMinBlocksPrimal = 3, kMinBlocksDual = 4;

// `q`: typed, host-folded parameters of the fast body; `d`, `P`: the full descriptors, touched only by the
// (out-of-line) generic fallback.  All by-value structs are __grid_constant__ so taking their address does not
// force a per-thread local copy.
template <class T, int NDIR, int VEC, bool NORMS>
__global__ void __launch_bounds__(kBlock, kMinBlocksPrimal) k_tv_primal(int algo, const __grid_constant__ PxbTvP<T> q,
                                                                        const __grid_constant__ pxb_grad_desc d,
                                                                        const __grid_constant__ pxb_pds_params P,
                                                                        const __grid_constant__ VoxMap m, T* __restrict__ xu,
                                                                        const T* __restrict__ z, T* __restrict__ x_out,
                                                                        T* __restrict__ w, double* __restrict__ norms) {
    const Vox v = vox_of_thread(m);
    double nrm[2] = {0.0, 0.0};
    if (v.ok) {
        if (algo == PXB_PD3O) pxb_tv_primal_vec<T, NDIR, VEC, PXB_PD3O, NORMS>(q, d, P, xu, z, x_out, w, nrm, v.b, v.i0, v.i1, v.i2 * VEC);
        else pxb_tv_primal_vec<T, NDIR, VEC, PXB_CV, NORMS>(q, d, P, xu, z, x_out, w, nrm, v.b, v.i0, v.i1, v.i2 * VEC);
    }
    if (NORMS) block_accumulate(nrm[0], nrm[1], v.b, v.ok, norms);
}

template <class T, int NDIR, int VEC, bool NORMS>
__global__ void __launch_bounds__(kBlock, kMinBlocksDual) k_tv_dual(const __grid_constant__ PxbTvP<T> q, const __grid_constant__ pxb_grad_desc d,
                                                                    const __grid_constant__ pxb_pds_params P,
                                                                    const __grid_constant__ VoxMap m, const T* __restrict__ w,
                                                                    T* __restrict__ z, double* __restrict__ norms) {
    const Vox v = vox_of_thread(m);
    double nrm[2] = {0.0, 0.0};
    if (v.ok) pxb_tv_dual_vec<T, NDIR, VEC, NORMS>(q, d, P, w, z, nrm, v.b, v.i0, v.i1, v.i2 * VEC);
    if (NORMS) block_accumulate(nrm[0], nrm[1], v.b, v.ok, norms);
}

template <class T, int NDIR, int VEC, bool ADJ>
__global__ void __launch_bounds__(kBlock, 4) k_tv_grad(const __grid_constant__ PxbTvP<T> q, const __grid_constant__ pxb_grad_desc d,
                                                       const __grid_constant__ VoxMap m, const T* __restrict__ in, T* __restrict__ out) {
    const Vox v = vox_of_thread(m);
    if (!v.ok) return;
    if (ADJ) pxb_tv_grad_adjoint_vec<T, NDIR, VEC>(q, d, in, out, v.b, v.i0, v.i1, v.i2 * VEC);
    else pxb_tv_grad_apply_vec<T, NDIR, VEC>(q, d, in, out, v.b, v.i0, v.i1, v.i2 * VEC);
}

// widest vector (<= 16 bytes) that divides the row length and matches every pointer's alignment
template <class T>
int pick_vec(int64_t n2, std::initializer_list<const void*> ptrs) {
    int vec = 16 / (int)sizeof(T);
    while (vec > 1) {
        bool ok = (n2 % vec) == 0;
        for (const void* p : ptrs)
            if (p && (reinterpret_cast<uintptr_t>(p) % (vec * sizeof(T))) != 0) ok = false;
        if (ok) break;
        vec >>= 1;
    }
    return vec;
}

bool make_map_vec(int64_t batch, const int64_t shape[3], int vec, VoxMap& m) {
    int64_t sh[3] = {shape[0], shape[1], shape[2] / vec};
    return make_map(batch, sh, m);
}

template <class T, int NDIR, int VEC>
void primal_launch(const VoxMap& m, cudaStream_t s, int algo, const pxb_grad_desc& d, const PxbTvCoef& cf, const pxb_pds_params& P,
                   void* xu, const void* z, void* x_out, void* w, double* norms) {
    PxbTvP<T> q;
    pxb_tv_prepare<T>(d, cf, P, q);
    if (norms) k_tv_primal<T, NDIR, VEC, true><<<grid_of(m), kBlock, 0, s>>>(algo, q, d, P, m, (T*)xu, (const T*)z, (T*)x_out, (T*)w, norms);
    else k_tv_primal<T, NDIR, VEC, false><<<grid_of(m), kBlock, 0, s>>>(algo, q, d, P, m, (T*)xu, (const T*)z, (T*)x_out, (T*)w, norms);
}

template <class T, int NDIR, int VEC>
void dual_launch(const VoxMap& m, cudaStream_t s, const pxb_grad_desc& d, const PxbTvCoef& cf, const pxb_pds_params& P, const void* w,
                 void* z, double* norms) {
    PxbTvP<T> q;
    pxb_tv_prepare<T>(d, cf, P, q);
    if (norms) k_tv_dual<T, NDIR, VEC, true><<<grid_of(m), kBlock, 0, s>>>(q, d, P, m, (const T*)w, (T*)z, norms);
    else k_tv_dual<T, NDIR, VEC, false><<<grid_of(m), kBlock, 0, s>>>(q, d, P, m, (const T*)w, (T*)z, norms);
}

template <class T, int NDIR, class... A>
void primal_vec(int vec, A&&... a) {
    if constexpr (sizeof(T) == 4) {
        if (vec == 4) return primal_launch<T, NDIR, 4>(a...);
    }
    if (vec == 2) return primal_launch<T, NDIR, 2>(a...);
    primal_launch<T, NDIR, 1>(a...);
}

template <class T, int NDIR, class... A>
void dual_vec(int vec, A&&... a) {
    if constexpr (sizeof(T) == 4) {
        if (vec == 4) return dual_launch<T, NDIR, 4>(a...);
    }
    if (vec == 2) return dual_launch<T, NDIR, 2>(a...);
    dual_launch<T, NDIR, 1>(a...);
}

template <class T, class... A>
void primal_dir(int ndir, int vec, A&&... a) {
    if (ndir == 3) primal_vec<T, 3>(vec, a...);
    else if (ndir == 2) primal_vec<T, 2>(vec, a...);
    else primal_vec<T, 1>(vec, a...);
}

template <class T, class... A>
void dual_dir(int ndir, int vec, A&&... a) {
    if (ndir == 3) dual_vec<T, 3>(vec, a...);
    else if (ndir == 2) dual_vec<T, 2>(vec, a...);
    else dual_vec<T, 1>(vec, a...);
}

}  // namespace

namespace {

template <class T, int NDIR, bool ADJ>
void grad_vec(int vec, const VoxMap& m, cudaStream_t s, const pxb_grad_desc& d, const PxbTvCoef& cf, const void* in, void* out) {
    pxb_pds_params P{};
    P.hkind = PXB_DUAL_NONE;
    PxbTvP<T> q;
    pxb_tv_prepare<T>(d, cf, P, q);
    if constexpr (sizeof(T) == 4) {
        if (vec == 4) { k_tv_grad<T, NDIR, 4, ADJ><<<grid_of(m), kBlock, 0, s>>>(q, d, m, (const T*)in, (T*)out); return; }
    }
    if (vec == 2) k_tv_grad<T, NDIR, 2, ADJ><<<grid_of(m), kBlock, 0, s>>>(q, d, m, (const T*)in, (T*)out);
    else k_tv_grad<T, NDIR, 1, ADJ><<<grid_of(m), kBlock, 0, s>>>(q, d, m, (const T*)in, (T*)out);
}

template <class T, bool ADJ>
void grad_dir(int ndir, int vec, const VoxMap& m, cudaStream_t s, const pxb_grad_desc& d, const PxbTvCoef& cf, const void* in, void* out) {
    if (ndir == 3) grad_vec<T, 3, ADJ>(vec, m, s, d, cf, in, out);
    else if (ndir == 2) grad_vec<T, 2, ADJ>(vec, m, s, d, cf, in, out);
    else grad_vec<T, 1, ADJ>(vec, m, s, d, cf, in, out);
}

}  // namespace

// Gradient.apply / adjoint through the vectorised bodies; false when the descriptor is not a first-order stack
bool pxb_tv_try_grad(const pxb_grad_desc* K, bool adjoint, const void* in, void* out, cudaStream_t s, int* rc) {
    PxbTvCoef cf;
    if (!pxb_tv_fast_coefs(*K, cf)) return false;
    const bool f32 = K->dtype == PXB_F32;
    const int vec = f32 ? pick_vec<float>(K->shape[2], {in, out}) : pick_vec<double>(K->shape[2], {in, out});
    VoxMap m;
    if (!make_map_vec(K->batch, K->shape, vec, m)) {
        *rc = pxb_fail(PXB_ENOSUP, "pxb_gradient: grid too large");
        return true;
    }
    if (f32) { if (adjoint) grad_dir<float, true>(K->ndir, vec, m, s, *K, cf, in, out); else grad_dir<float, false>(K->ndir, vec, m, s, *K, cf, in, out); }
    else { if (adjoint) grad_dir<double, true>(K->ndir, vec, m, s, *K, cf, in, out); else grad_dir<double, false>(K->ndir, vec, m, s, *K, cf, in, out); }
    pxb_count_launch();
    cudaError_t e = cudaGetLastError();
    *rc = e == cudaSuccess ? 0 : pxb_fail(PXB_ECUDA, "pxb_gradient: %s", cudaGetErrorString(e));
    return true;
}

bool pxb_tv_try_primal(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, void* xu, const void* z, void* x_out, void* w,
                       double* norms, cudaStream_t s, int* rc) {
    PxbTvCoef cf;
    if (!pxb_tv_fast_coefs(*K, cf)) return false;
    const bool f32 = K->dtype == PXB_F32;
    const void* sh = (p->f.kind == PXB_F_SQL2 && p->f.shift_period > 1) ? p->f.shift : nullptr;
    const void* ga = p->f.kind == PXB_F_GRADARR ? p->f.garr : nullptr;
    const int vec = f32 ? pick_vec<float>(K->shape[2], {xu, z, x_out, w, sh, ga}) : pick_vec<double>(K->shape[2], {xu, z, x_out, w, sh, ga});
    VoxMap m;
    if (!make_map_vec(K->batch, K->shape, vec, m)) {
        *rc = pxb_fail(PXB_ENOSUP, "pxb_pds_primal: grid too large");
        return true;
    }
    if (f32) primal_dir<float>(K->ndir, vec, m, s, algo, *K, cf, *p, xu, z, x_out, w, norms);
    else primal_dir<double>(K->ndir, vec, m, s, algo, *K, cf, *p, xu, z, x_out, w, norms);
    pxb_count_launch();
    cudaError_t e = cudaGetLastError();
    *rc = e == cudaSuccess ? 0 : pxb_fail(PXB_ECUDA, "pxb_pds_primal: %s", cudaGetErrorString(e));
    return true;
}

bool pxb_tv_try_dual(const pxb_grad_desc* K, const pxb_pds_params* p, const void* w, void* z, double* norms, cudaStream_t s, int* rc) {
    PxbTvCoef cf;
    if (!pxb_tv_fast_coefs(*K, cf)) return false;
    const bool f32 = K->dtype == PXB_F32;
    const int vec = f32 ? pick_vec<float>(K->shape[2], {w, z}) : pick_vec<double>(K->shape[2], {w, z});
    VoxMap m;
    if (!make_map_vec(K->batch, K->shape, vec, m)) {
        *rc = pxb_fail(PXB_ENOSUP, "pxb_pds_dual: grid too large");
        return true;
    }
    if (f32) dual_dir<float>(K->ndir, vec, m, s, *K, cf, *p, w, z, norms);
    else dual_dir<double>(K->ndir, vec, m, s, *K, cf, *p, w, z, norms);
    pxb_count_launch();
    cudaError_t e = cudaGetLastError();
    *rc = e == cudaSuccess ? 0 : pxb_fail(PXB_ECUDA, "pxb_pds_dual: %s", cudaGetErrorString(e));
    return true;
}
