// pxb_kernels.cu -- CUDA kernels (sm_100a) + C ABI of libpyxu_b200.so.
//
// Every kernel on this path is HBM-bound integer-free streaming work (a handful of FMAs per byte
// moved), so the design rules are: one pass per voxel, coalesced accesses along the fastest axis,
// neighbour taps served by L1/L2 (a z-plane of 1024^2 fp32 is 4 MiB, the L2 is 126 MB), grid sized
// from the volume, fp64 accumulation of the stopping-criterion norms with one atomic per block.
// Tensor cores are deliberately unused: nothing here is a dense contraction.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "pxb_launch.cuh"
#include "pxb_tv_fast.cuh"

namespace {
thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};
}  // namespace

int pxb_fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
void pxb_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
#define fail pxb_fail

namespace {

// ------------------------------------------------------------------------------------------
// Kernels
// ------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(kBlock) k_stencil(pxb_stencil_desc d, VoxMap m, int adjoint, const T* __restrict__ in, T* __restrict__ out) {
    const Vox v = vox_of_thread(m);
    if (!v.ok) return;
    const PxbGeom g = pxb_geom(d.shape);
    pxb_body_stencil<T>(d, g, adjoint != 0, in, out, v.b, v.i0, v.i1, v.i2);
}

template <class T>
__global__ void __launch_bounds__(kBlock) k_grad_apply(pxb_grad_desc d, VoxMap m, const T* __restrict__ x, T* __restrict__ z) {
    const Vox v = vox_of_thread(m);
    if (!v.ok) return;
    const PxbGeom g = pxb_geom(d.shape);
    pxb_body_grad_apply<T>(d, g, x, z, v.b, v.i0, v.i1, v.i2);
}

template <class T>
__global__ void __launch_bounds__(kBlock) k_grad_adjoint(pxb_grad_desc d, VoxMap m, const T* __restrict__ z, T* __restrict__ x) {
    const Vox v = vox_of_thread(m);
    if (!v.ok) return;
    const PxbGeom g = pxb_geom(d.shape);
    pxb_body_grad_adjoint<T>(d, g, z, x, v.b, v.i0, v.i1, v.i2);
}

template <class T>
__global__ void __launch_bounds__(kBlock) k_pds_primal(int algo, pxb_grad_desc d, pxb_pds_params P, VoxMap m, T* __restrict__ xu,
                                                       const T* __restrict__ z, const T* __restrict__ ktz, T* __restrict__ x_out,
                                                       T* __restrict__ w, double* __restrict__ norms) {
    const Vox v = vox_of_thread(m);
    const PxbGeom g = pxb_geom(d.shape);
    double a0 = 0.0, a1 = 0.0;
    if (v.ok) pxb_body_primal<T>(algo, d, g, P, xu, z, ktz, x_out, w, norms != nullptr, a0, a1, v.b, v.i0, v.i1, v.i2);
    if (norms) block_accumulate(a0, a1, v.b, v.ok, norms);
}

template <class T>
__global__ void __launch_bounds__(kBlock) k_pds_dual(pxb_grad_desc d, pxb_pds_params P, VoxMap m, const T* __restrict__ w,
                                                     T* __restrict__ z, double* __restrict__ norms) {
    const Vox v = vox_of_thread(m);
    const PxbGeom g = pxb_geom(d.shape);
    double a0 = 0.0, a1 = 0.0;
    if (v.ok) pxb_body_dual<T>(d, g, P, w, z, norms != nullptr, a0, a1, v.b, v.i0, v.i1, v.i2);
    if (norms) block_accumulate(a0, a1, v.b, v.ok, norms);
}

// (outer, group, inner) kernels: one thread per (outer, inner) pair, `inner` fastest.
template <class T>
__global__ void __launch_bounds__(kBlock) k_dual_update(int kind, int64_t outer, int64_t group, int64_t inner, T lam, T sigma, T rho,
                                                        T* __restrict__ z, const T* __restrict__ t, double* __restrict__ norms) {
    const int64_t per = (inner + kBlock - 1) / kBlock;  // blocks per outer row
    const int64_t o = blockIdx.x / per;
    const int64_t i = (blockIdx.x % per) * kBlock + threadIdx.x;
    const bool ok = i < inner;
    double a0 = 0.0, a1 = 0.0;
    if (ok) pxb_body_dual_update<T>(kind, group, inner, lam, sigma, rho, z, t, norms != nullptr, a0, a1, o, i);
    if (norms) block_accumulate(a0, a1, o, ok, norms);
}

template <class T>
__global__ void __launch_bounds__(kBlock) k_prox_l21(int64_t outer, int64_t group, int64_t inner, T lam, T tau, const T* __restrict__ x,
                                                     T* __restrict__ out) {
    const int64_t per = (inner + kBlock - 1) / kBlock;
    const int64_t o = blockIdx.x / per;
    const int64_t i = (blockIdx.x % per) * kBlock + threadIdx.x;
    if (i < inner) pxb_body_prox_l21<T>(group, inner, lam, tau, x, out, o, i);
}

template <class T, bool PROX>
__global__ void __launch_bounds__(kBlock) k_lincomb(pxb_prox_spec g, T tau, int64_t n, T* out, T a, const T* x, T b, const T* y,
                                                    int64_t ny, T c, const T* z, int64_t nz) {
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
        T v = pxb_lincomb_at<T>(a, x, b, y, ny, c, z, nz, i);
        if (PROX) v = pxb_prox_eval<T>(g.kind, T(g.p0), T(g.p1), v, tau);
        out[i] = v;
    }
}

// 128-bit forms of the elementwise kernels (n, the broadcast periods and every base address multiples of the vector):
// the scalar forms reached 4.1-4.6 TB/s, i.e. 0.63-0.70 of the measured copy bandwidth.
template <class T, int VEC, bool PROX>
__global__ void __launch_bounds__(kBlock) k_lincomb_vec(pxb_prox_spec g, T tau, int64_t nv, T* out, T a, const T* x, T b, const T* y,
                                                        int64_t ny, T c, const T* z, int64_t nz) {
    for (int64_t iv = (int64_t)blockIdx.x * kBlock + threadIdx.x; iv < nv; iv += (int64_t)gridDim.x * kBlock) {
        const int64_t i = iv * VEC;
        PxbVec<T, VEC> v = pxb_vload<T, VEC>(x + i);
        for (int j = 0; j < VEC; ++j) v.v[j] *= a;
        if (y) {
            const PxbVec<T, VEC> t = pxb_vload<T, VEC>(y + (ny ? i % ny : i));
            for (int j = 0; j < VEC; ++j) v.v[j] += b * t.v[j];
        }
        if (z) {
            const PxbVec<T, VEC> t = pxb_vload<T, VEC>(z + (nz ? i % nz : i));
            for (int j = 0; j < VEC; ++j) v.v[j] += c * t.v[j];
        }
        if (PROX)
            for (int j = 0; j < VEC; ++j) v.v[j] = pxb_prox_eval<T>(g.kind, T(g.p0), T(g.p1), v.v[j], tau);
        pxb_vstore<T, VEC>(out + i, v);
    }
}

template <class T, int VEC>
__global__ void __launch_bounds__(kBlock) k_sqnorms_vec(int64_t rows, int64_t n, const T* __restrict__ x, const T* __restrict__ y,
                                                        double* __restrict__ out, int64_t per) {
    const int64_t r = blockIdx.x / per;
    const int64_t chunk = blockIdx.x % per;
    double a0 = 0.0, a1 = 0.0;
    for (int64_t i = (chunk * kBlock + threadIdx.x) * VEC; i < n; i += per * kBlock * VEC) {
        const PxbVec<T, VEC> xv = pxb_vload<T, VEC>(x + r * n + i);
        if (y) {
            const PxbVec<T, VEC> yv = pxb_vload<T, VEC>(y + r * n + i);
            for (int j = 0; j < VEC; ++j) {
                const double d = (double)xv.v[j] - (double)yv.v[j];
                a0 += d * d;
                a1 += (double)yv.v[j] * (double)yv.v[j];
            }
        } else {
            for (int j = 0; j < VEC; ++j) a0 += (double)xv.v[j] * (double)xv.v[j];
        }
    }
    block_accumulate(a0, a1, r, true, out);
}

// group soft-threshold out = x * (1 - t / max(||x_group||, t)) with VEC consecutive `inner` samples per thread, GROUP <= 3
template <class T, int VEC>
__global__ void __launch_bounds__(kBlock) k_prox_l21_vec(int64_t outer, int group, int64_t inner, T lam, T tau, const T* __restrict__ x,
                                                         T* __restrict__ out) {
    const int64_t innerv = inner / VEC;
    const int64_t per = (innerv + kBlock - 1) / kBlock;
    const int64_t o = blockIdx.x / per;
    const int64_t iv = (blockIdx.x % per) * kBlock + threadIdx.x;
    if (iv >= innerv) return;
    const int64_t base = o * group * inner + iv * VEC;
    const T t = tau * lam;
    PxbVec<T, VEC> v[3];
    T nn[VEC];
    for (int j = 0; j < VEC; ++j) nn[j] = T(0);
    for (int k = 0; k < group; ++k) {
        v[k] = pxb_vload<T, VEC>(x + base + k * inner);
        for (int j = 0; j < VEC; ++j) nn[j] += v[k].v[j] * v[k].v[j];
    }
    T sc[VEC];
    for (int j = 0; j < VEC; ++j) {
        const T nrm = sqrt(nn[j]);
        sc[j] = T(1) - t / (nrm > t ? nrm : t);
    }
    for (int k = 0; k < group; ++k) {
        for (int j = 0; j < VEC; ++j) v[k].v[j] *= sc[j];
        pxb_vstore<T, VEC>(out + base + k * inner, v[k]);
    }
}

// dual update z <- (1-rho) z + rho prox_{sigma h*}(z + sigma t) with VEC consecutive `inner` samples per thread, GROUP <= 3
template <class T, int VEC>
__global__ void __launch_bounds__(kBlock) k_dual_update_vec(int kind, int64_t outer, int group, int64_t inner, T lam, T sigma, T rho,
                                                            T* __restrict__ z, const T* __restrict__ t, double* __restrict__ norms) {
    const int64_t innerv = inner / VEC;
    const int64_t per = (innerv + kBlock - 1) / kBlock;
    const int64_t o = blockIdx.x / per;
    const int64_t iv = (blockIdx.x % per) * kBlock + threadIdx.x;
    const bool ok = iv < innerv;
    double a0 = 0.0, a1 = 0.0;
    if (ok) {
        const int64_t base = o * group * inner + iv * VEC;
        PxbVec<T, VEC> zo[3];
        T p[VEC][PXB_MAX_DIRS];
        for (int k = 0; k < group; ++k) {
            zo[k] = pxb_vload<T, VEC>(z + base + k * inner);
            const PxbVec<T, VEC> tv = pxb_vload<T, VEC>(t + base + k * inner);
            for (int j = 0; j < VEC; ++j) p[j][k] = zo[k].v[j] + sigma * tv.v[j];
        }
        for (int j = 0; j < VEC; ++j) pxb_dual_prox_group<T>(kind, group, lam, sigma, p[j]);
        for (int k = 0; k < group; ++k) {
            PxbVec<T, VEC> zn;
            for (int j = 0; j < VEC; ++j) {
                zn.v[j] = (T(1) - rho) * zo[k].v[j] + rho * p[j][k];
                if (norms) {
                    const double dd = (double)zn.v[j] - (double)zo[k].v[j];
                    a0 += dd * dd;
                    a1 += (double)zo[k].v[j] * (double)zo[k].v[j];
                }
            }
            pxb_vstore<T, VEC>(z + base + k * inner, zn);
        }
    }
    if (norms) block_accumulate(a0, a1, o, ok, norms);
}

template <class T>
__global__ void __launch_bounds__(kBlock) k_sqnorms(int64_t rows, int64_t n, const T* __restrict__ x, const T* __restrict__ y,
                                                    double* __restrict__ out, int64_t per) {
    const int64_t r = blockIdx.x / per;
    const int64_t chunk = blockIdx.x % per;
    double a0 = 0.0, a1 = 0.0;
    // each block strides over its row: chunk c covers i = c*kBlock + t, + per*kBlock, ...
    for (int64_t i = chunk * kBlock + threadIdx.x; i < n; i += per * kBlock) {
        const double xv = (double)x[r * n + i];
        if (y) {
            const double yv = (double)y[r * n + i];
            a0 += (xv - yv) * (xv - yv);
            a1 += yv * yv;
        } else {
            a0 += xv * xv;
        }
    }
    block_accumulate(a0, a1, r, true, out);
}

// ------------------------------------------------------------------------------------------
// Argument validation
// ------------------------------------------------------------------------------------------
int check_shape(const int64_t shape[3], int64_t batch, const char* who) {
    if (batch < 1) return fail(PXB_EINVAL, "%s: batch must be >= 1", who);
    for (int a = 0; a < 3; ++a)
        if (shape[a] < 1 || shape[a] > 0x3fffffff) return fail(PXB_EINVAL, "%s: shape[%d]=%lld out of range", who, a, (long long)shape[a]);
    return 0;
}

// pad-width limits of the reference (pad.py:217-229): a boundary coordinate must fold at most once.
int check_mode(int mode, int64_t n, int p, int open_lo, int open_hi, const char* who, int axis) {
    if (mode < PXB_CONSTANT || mode > PXB_EDGE) return fail(PXB_EINVAL, "%s: unknown mode %d on axis %d", who, mode, axis);
    if (open_lo || open_hi) return 0;  // slab / sub-range launch: the caller guarantees the halo covers the reach
    int64_t lim = 0x7fffffff;
    if (mode == PXB_WRAP || mode == PXB_SYMMETRIC) lim = n;
    if (mode == PXB_REFLECT) lim = n - 1;
    if (p > lim) return fail(PXB_EINVAL, "%s: pad width %d along axis %d is limited to %lld for this mode", who, p, axis, (long long)lim);
    return 0;
}

int check_slab(const pxb_slab& s, int64_t batch, int need, const char* who) {
    if (s.halo < 0) return fail(PXB_EINVAL, "%s: negative halo", who);
    if ((s.open_lo || s.open_hi) && s.halo < need) return fail(PXB_EINVAL, "%s: halo=%d planes < stencil reach %d", who, s.halo, need);
    if (s.halo > 0 && batch != 1) return fail(PXB_EINVAL, "%s: slab halos require batch == 1", who);
    if (s.plane_alloc < 0) return fail(PXB_EINVAL, "%s: negative plane_alloc", who);
    return 0;
}

int check_stencil(const pxb_stencil_desc* d, const void* in, const void* out, const char* who) {
    if (!d || !in || !out) return fail(PXB_EINVAL, "%s: null argument", who);
    if (in == out) return fail(PXB_EINVAL, "%s: in and out must not alias", who);
    if (d->dtype != PXB_F32 && d->dtype != PXB_F64) return fail(PXB_EINVAL, "%s: bad dtype %d", who, d->dtype);
    if (!d->coef) return fail(PXB_EINVAL, "%s: null coefficient pointer", who);
    if (int e = check_shape(d->shape, d->batch, who)) return e;
    for (int a = 0; a < 3; ++a) {
        if (d->ksize[a] < 1) return fail(PXB_EINVAL, "%s: ksize[%d] < 1", who, a);
        if (d->center[a] < 0 || d->center[a] >= d->ksize[a]) return fail(PXB_EINVAL, "%s: center[%d] outside kernel", who, a);
        const int p = d->mode[a] == PXB_CONSTANT ? 0 : d->ksize[a] - 1;
        if (int e = check_mode(d->mode[a], d->shape[a], p, a == 0 ? d->slab.open_lo : 0, a == 0 ? d->slab.open_hi : 0, who, a)) return e;
    }
    return check_slab(d->slab, d->batch, d->ksize[0] - 1, who);
}

int check_grad(const pxb_grad_desc* d, const char* who) {
    if (!d) return fail(PXB_EINVAL, "%s: null descriptor", who);
    if (d->dtype != PXB_F32 && d->dtype != PXB_F64) return fail(PXB_EINVAL, "%s: bad dtype %d", who, d->dtype);
    if (d->ndir < 1 || d->ndir > PXB_MAX_DIRS) return fail(PXB_ENOSUP, "%s: ndir=%d outside [1,%d]", who, d->ndir, PXB_MAX_DIRS);
    if (int e = check_shape(d->shape, d->batch, who)) return e;
    int reach0 = 0;
    for (int k = 0; k < d->ndir; ++k) {
        const int ax = d->axis[k];
        if (ax < 0 || ax > 2) return fail(PXB_EINVAL, "%s: axis[%d]=%d", who, k, ax);
        if (d->ntap[k] < 1 || d->ntap[k] > PXB_MAX_GTAP) return fail(PXB_ENOSUP, "%s: ntap[%d]=%d outside [1,%d]", who, k, d->ntap[k], PXB_MAX_GTAP);
        if (d->center[k] < 0 || d->center[k] >= d->ntap[k]) return fail(PXB_EINVAL, "%s: center[%d] outside kernel", who, k);
        const int p = d->mode[ax] == PXB_CONSTANT ? 0 : d->ntap[k] - 1;
        if (int e = check_mode(d->mode[ax], d->shape[ax], p, ax == 0 ? d->slab.open_lo : 0, ax == 0 ? d->slab.open_hi : 0, who, ax)) return e;
        if (ax == 0 && d->ntap[k] - 1 > reach0) reach0 = d->ntap[k] - 1;
    }
    return check_slab(d->slab, d->batch, reach0, who);
}

unsigned flat_grid(int64_t n) {
    int64_t b = (n + kBlock - 1) / kBlock;
    const int64_t cap = 148LL * 32;  // grid-stride: a few waves of the 148 SMs
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

// Pad / Pad^T over the two trailing axes (pxb_pad2d / pxb_pad2d_adjoint): one thread per written sample
template <class T>
__global__ void __launch_bounds__(kBlock) k_pad2d(const __grid_constant__ pxb_pad2d_desc d, const __grid_constant__ VoxMap m, const T* __restrict__ in,
                                                  T* __restrict__ ext) {
    const Vox v = vox_of_thread(m);
    if (!v.ok) return;
    ext[(v.b * d.ext_shape[0] + v.i1) * d.ext_shape[1] + v.i2] = pxb_pad2d_at<T>(d, in, v.b, v.i1, v.i2);
}
template <class T>
__global__ void __launch_bounds__(kBlock) k_pad2d_adjoint(const __grid_constant__ pxb_pad2d_desc d, const __grid_constant__ VoxMap m,
                                                          const T* __restrict__ ext, T* __restrict__ out, T alpha, T beta, const T* __restrict__ add,
                                                          int64_t add_period) {
    const Vox v = vox_of_thread(m);
    if (!v.ok) return;
    const int64_t lin = (v.b * d.shape[0] + v.i1) * d.shape[1] + v.i2;
    T o = alpha * pxb_pad2d_adj_at<T>(d, ext, v.b, v.i1, v.i2);
    if (add) o += beta * add[add_period > 0 ? lin % add_period : lin];
    out[lin] = o;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
static std::atomic<int> g_iter_path{-1};
int pxb_iter_path() {
    int v = g_iter_path.load();
    if (v < 0) {  // first use: PXB_TV_ITER=direct|tma overrides the automatic choice
        const char* e = getenv("PXB_TV_ITER");
        v = (e && !strcmp(e, "direct")) ? 1 : (e && !strcmp(e, "tma")) ? 2 : 0;
        g_iter_path.store(v);
    }
    return v;
}

// Default of the folding-mode instances of the single-kernel iteration (see pxb_set_iter_modes): on.
#ifndef PXB_ITER_MODES_DEFAULT
#define PXB_ITER_MODES_DEFAULT 1
#endif
static std::atomic<int> g_iter_modes{-1};
int pxb_iter_modes() {
    int v = g_iter_modes.load();
    if (v < 0) {
        const char* e = getenv("PXB_TV_ITER_MODES");
        v = e ? (atoi(e) != 0 ? 1 : 0) : PXB_ITER_MODES_DEFAULT;
        g_iter_modes.store(v);
    }
    return v;
}

extern "C" {
int pxb_set_iter_modes(int on) {
    if (on < -1 || on > 1) return fail(PXB_EINVAL, "pxb_set_iter_modes: 0, 1 or -1 (back to the initial value)");
    g_iter_modes.store(on);
    return 0;
}

int pxb_set_iter_path(int path) {
    if (path < 0 || path > 2) return fail(PXB_EINVAL, "pxb_set_iter_path: 0 (auto), 1 (direct loads) or 2 (TMA)");
    g_iter_path.store(path);
    return 0;
}

int pxb_abi_version(void) { return PXB_ABI_VERSION; }

int pxb_enable_peer_access(int peer_device) {
    int me = -1;
    cudaError_t e = cudaGetDevice(&me);
    if (e != cudaSuccess) return pxb_fail(PXB_ECUDA, "pxb_enable_peer_access: %s", cudaGetErrorString(e));
    if (me == peer_device) return 0;
    int can = 0;
    e = cudaDeviceCanAccessPeer(&can, me, peer_device);
    if (e != cudaSuccess || !can) {
        (void)cudaGetLastError();
        return pxb_fail(PXB_ENOSUP, "pxb_enable_peer_access: device %d cannot access device %d", me, peer_device);
    }
    e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); return 0; }
    if (e != cudaSuccess) return pxb_fail(PXB_ECUDA, "pxb_enable_peer_access: %s", cudaGetErrorString(e));
    return 0;
}
const char* pxb_last_error(void) { return g_err.c_str(); }
int64_t pxb_launch_count(void) { return g_launches.load(); }

static int stencil_launch(const pxb_stencil_desc* d, const void* in, void* out, void* stream, int adjoint, const char* who) {
    if (int e = check_stencil(d, in, out, who)) return e;
    VoxMap m;
    if (!make_map(d->batch, d->shape, m)) return fail(PXB_ENOSUP, "%s: grid too large", who);
    cudaStream_t s = (cudaStream_t)stream;
    if (d->dtype == PXB_F32) k_stencil<float><<<grid_of(m), kBlock, 0, s>>>(*d, m, adjoint, (const float*)in, (float*)out);
    else k_stencil<double><<<grid_of(m), kBlock, 0, s>>>(*d, m, adjoint, (const double*)in, (double*)out);
    PXB_CHECK_LAUNCH(who);
    return 0;
}

int pxb_stencil_apply(const pxb_stencil_desc* d, const void* in, void* out, void* stream) {
    return stencil_launch(d, in, out, stream, 0, "pxb_stencil_apply");
}
int pxb_stencil_adjoint(const pxb_stencil_desc* d, const void* in, void* out, void* stream) {
    return stencil_launch(d, in, out, stream, 1, "pxb_stencil_adjoint");
}

int pxb_stencil2d_apply(const pxb_stencil2d* d, const void* in, void* out, void* stream) {
    const char* who = "pxb_stencil2d_apply";
    if (!d || !in || !out) return fail(PXB_EINVAL, "%s: null argument", who);
    if (d->dtype != PXB_F32 && d->dtype != PXB_F64) return fail(PXB_EINVAL, "%s: bad dtype %d", who, d->dtype);
    if (in == out) return fail(PXB_EINVAL, "%s: in and out must not alias", who);
    if (d->nimg < 1 || d->shape[0] < 1 || d->shape[1] < 1) return fail(PXB_EINVAL, "%s: empty array", who);
    if (d->dense && !d->coef) return fail(PXB_EINVAL, "%s: dense kernel without coefficients", who);
    cudaError_t err = cudaSuccess;
    if (int why = pxb_stencil2d_try(d, in, out, (cudaStream_t)stream, &err)) return fail(PXB_ENOSUP, "%s: outside the tiled kernel's envelope (reason %d)", who, why);
    pxb_count_launch();
    if (err != cudaSuccess) return fail(PXB_ECUDA, "%s: %s", who, cudaGetErrorString(err));
    return 0;
}

static int pad2d_check(const pxb_pad2d_desc* d, const char* who) {
    if (!d) return fail(PXB_EINVAL, "%s: null descriptor", who);
    if (d->dtype != PXB_F32 && d->dtype != PXB_F64) return fail(PXB_EINVAL, "%s: bad dtype %d", who, d->dtype);
    if (d->nimg < 1 || d->shape[0] < 1 || d->shape[1] < 1) return fail(PXB_EINVAL, "%s: empty array", who);
    for (int a = 0; a < 2; ++a) {
        if (d->lo[a] < 0 || d->hi[a] < 0 || d->org[a] < d->lo[a] || d->ext_shape[a] < d->org[a] + d->shape[a] + d->hi[a])
            return fail(PXB_EINVAL, "%s: padded extent too small along axis %d", who, a);
        if (d->mode[a] < PXB_CONSTANT || d->mode[a] > PXB_EDGE) return fail(PXB_EINVAL, "%s: bad mode %d", who, d->mode[a]);
        // a coordinate folds at most once (pad.py:217-229)
        const int64_t n = d->shape[a], w = d->lo[a] > d->hi[a] ? d->lo[a] : d->hi[a];
        const int64_t lim = d->mode[a] == PXB_REFLECT ? n - 1 : ((d->mode[a] == PXB_WRAP || d->mode[a] == PXB_SYMMETRIC) ? n : w);
        if (w > lim) return fail(PXB_EINVAL, "%s: pad width %lld along axis %d exceeds the limit %lld of the mode", who, (long long)w, a, (long long)lim);
    }
    return 0;
}

int pxb_pad2d(const pxb_pad2d_desc* d, const void* in, void* ext, void* stream) {
    const char* who = "pxb_pad2d";
    if (int rc = pad2d_check(d, who)) return rc;
    if (!in || !ext || in == ext) return fail(PXB_EINVAL, "%s: null or aliased arrays", who);
    VoxMap m;
    const int64_t sh[3] = {1, d->ext_shape[0], d->ext_shape[1]};
    if (!make_map(d->nimg, sh, m)) return fail(PXB_ENOSUP, "%s: grid too large", who);
    cudaStream_t s = (cudaStream_t)stream;
    if (d->dtype == PXB_F32) k_pad2d<float><<<grid_of(m), kBlock, 0, s>>>(*d, m, (const float*)in, (float*)ext);
    else k_pad2d<double><<<grid_of(m), kBlock, 0, s>>>(*d, m, (const double*)in, (double*)ext);
    PXB_CHECK_LAUNCH(who);
    return 0;
}

int pxb_pad2d_adjoint(const pxb_pad2d_desc* d, const void* ext, void* out, double alpha, double beta, const void* add, int64_t add_period,
                      void* stream) {
    const char* who = "pxb_pad2d_adjoint";
    if (int rc = pad2d_check(d, who)) return rc;
    if (!ext || !out || ext == out) return fail(PXB_EINVAL, "%s: null or aliased arrays", who);
    VoxMap m;
    const int64_t sh[3] = {1, d->shape[0], d->shape[1]};
    if (!make_map(d->nimg, sh, m)) return fail(PXB_ENOSUP, "%s: grid too large", who);
    if (add && add_period >= d->nimg * d->shape[0] * d->shape[1]) add_period = 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (d->dtype == PXB_F32) k_pad2d_adjoint<float><<<grid_of(m), kBlock, 0, s>>>(*d, m, (const float*)ext, (float*)out, (float)alpha, (float)beta, (const float*)add, add_period);
    else k_pad2d_adjoint<double><<<grid_of(m), kBlock, 0, s>>>(*d, m, (const double*)ext, (double*)out, alpha, beta, (const double*)add, add_period);
    PXB_CHECK_LAUNCH(who);
    return 0;
}

int pxb_stencil2d_fista(const pxb_stencil2d* d, const pxb_fista_step* f, int which, void* out, void* stream) {
    const char* who = "pxb_stencil2d_fista";
    if (!d || !f || !out || !f->x) return fail(PXB_EINVAL, "%s: null argument", who);
    if (d->dtype != PXB_F32 && d->dtype != PXB_F64) return fail(PXB_EINVAL, "%s: bad dtype %d", who, d->dtype);
    if (which != 0 && which != 1) return fail(PXB_EINVAL, "%s: which must be 0 (residual) or 1 (prox step)", who);
    if (f->a != 0.0 && !f->x_prev) return fail(PXB_EINVAL, "%s: momentum needs x_prev", who);
    if (which == 1 && !f->r) return fail(PXB_EINVAL, "%s: the prox step needs r", who);
    if (which == 1 && (f->g.kind < PXB_PROX_NONE || f->g.kind > PXB_PROX_SQL2)) return fail(PXB_EINVAL, "%s: bad g kind", who);
    if (which == 1 && (out == f->x || out == f->r)) return fail(PXB_EINVAL, "%s: out may alias x_prev only", who);
    if (which == 0 && (out == f->x || out == f->x_prev)) return fail(PXB_EINVAL, "%s: out must not alias the inputs", who);
    if (d->dense && !d->coef) return fail(PXB_EINVAL, "%s: dense kernel without coefficients", who);
    cudaError_t err = cudaSuccess;
    if (int why = pxb_stencil2d_fista_try(d, f, which, out, (cudaStream_t)stream, &err)) return fail(PXB_ENOSUP, "%s: outside the tiled kernel's envelope (reason %d)", who, why);
    pxb_count_launch();
    if (err != cudaSuccess) return fail(PXB_ECUDA, "%s: %s", who, cudaGetErrorString(err));
    return 0;
}

int pxb_gradient_apply(const pxb_grad_desc* d, const void* x, void* z, void* stream) {
    const char* who = "pxb_gradient_apply";
    if (int e = check_grad(d, who)) return e;
    if (!x || !z || x == z) return fail(PXB_EINVAL, "%s: null or aliased arrays", who);
    VoxMap m;
    cudaStream_t s = (cudaStream_t)stream;
    {
        int rc = 0;  // first-order stacks: 128-bit vectorised bodies (pxb_tv_kernels.cu)
        if (pxb_tv_try_grad(d, false, x, z, s, &rc)) return rc;
    }
    if (!make_map(d->batch, d->shape, m)) return fail(PXB_ENOSUP, "%s: grid too large", who);
    if (d->dtype == PXB_F32) k_grad_apply<float><<<grid_of(m), kBlock, 0, s>>>(*d, m, (const float*)x, (float*)z);
    else k_grad_apply<double><<<grid_of(m), kBlock, 0, s>>>(*d, m, (const double*)x, (double*)z);
    PXB_CHECK_LAUNCH(who);
    return 0;
}

int pxb_gradient_adjoint(const pxb_grad_desc* d, const void* z, void* x, void* stream) {
    const char* who = "pxb_gradient_adjoint";
    if (int e = check_grad(d, who)) return e;
    if (!x || !z || x == z) return fail(PXB_EINVAL, "%s: null or aliased arrays", who);
    VoxMap m;
    cudaStream_t s = (cudaStream_t)stream;
    {
        int rc = 0;
        if (pxb_tv_try_grad(d, true, z, x, s, &rc)) return rc;
    }
    if (!make_map(d->batch, d->shape, m)) return fail(PXB_ENOSUP, "%s: grid too large", who);
    if (d->dtype == PXB_F32) k_grad_adjoint<float><<<grid_of(m), kBlock, 0, s>>>(*d, m, (const float*)z, (float*)x);
    else k_grad_adjoint<double><<<grid_of(m), kBlock, 0, s>>>(*d, m, (const double*)z, (double*)x);
    PXB_CHECK_LAUNCH(who);
    return 0;
}

static int lincomb_launch(bool prox, int dtype, const pxb_prox_spec* g, double tau, int64_t n, void* out, double a, const void* x,
                          double b, const void* y, int64_t ny, double c, const void* z, int64_t nz, void* stream, const char* who) {
    if (dtype != PXB_F32 && dtype != PXB_F64) return fail(PXB_EINVAL, "%s: bad dtype %d", who, dtype);
    if (n < 0 || !out || !x) return fail(PXB_EINVAL, "%s: null array or negative size", who);
    if (ny < 0 || nz < 0) return fail(PXB_EINVAL, "%s: negative broadcast period", who);
    if (prox && (!g || g->kind < PXB_PROX_NONE || g->kind > PXB_PROX_SQL2)) return fail(PXB_EINVAL, "%s: bad prox spec", who);
    if (n == 0) return 0;
    pxb_prox_spec gs = prox ? *g : pxb_prox_spec{PXB_PROX_NONE, 0, 0.0, 0.0};
    if (ny >= n) ny = 0;
    if (nz >= n) nz = 0;
    cudaStream_t s = (cudaStream_t)stream;
    {   // 128-bit form when sizes, periods and addresses allow
        const int vec = dtype == PXB_F32 ? 4 : 2;
        const uintptr_t bits = reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(z);
        if (n % vec == 0 && ny % vec == 0 && nz % vec == 0 && (bits & 15u) == 0) {
            const int64_t nv = n / vec;
            const unsigned gridv = flat_grid(nv);
            if (dtype == PXB_F32) {
                if (prox) k_lincomb_vec<float, 4, true><<<gridv, kBlock, 0, s>>>(gs, (float)tau, nv, (float*)out, (float)a, (const float*)x, (float)b, (const float*)y, ny, (float)c, (const float*)z, nz);
                else k_lincomb_vec<float, 4, false><<<gridv, kBlock, 0, s>>>(gs, (float)tau, nv, (float*)out, (float)a, (const float*)x, (float)b, (const float*)y, ny, (float)c, (const float*)z, nz);
            } else {
                if (prox) k_lincomb_vec<double, 2, true><<<gridv, kBlock, 0, s>>>(gs, tau, nv, (double*)out, a, (const double*)x, b, (const double*)y, ny, c, (const double*)z, nz);
                else k_lincomb_vec<double, 2, false><<<gridv, kBlock, 0, s>>>(gs, tau, nv, (double*)out, a, (const double*)x, b, (const double*)y, ny, c, (const double*)z, nz);
            }
            PXB_CHECK_LAUNCH(who);
            return 0;
        }
    }
    const unsigned grid = flat_grid(n);
    if (dtype == PXB_F32) {
        if (prox) k_lincomb<float, true><<<grid, kBlock, 0, s>>>(gs, (float)tau, n, (float*)out, (float)a, (const float*)x, (float)b, (const float*)y, ny, (float)c, (const float*)z, nz);
        else k_lincomb<float, false><<<grid, kBlock, 0, s>>>(gs, (float)tau, n, (float*)out, (float)a, (const float*)x, (float)b, (const float*)y, ny, (float)c, (const float*)z, nz);
    } else {
        if (prox) k_lincomb<double, true><<<grid, kBlock, 0, s>>>(gs, tau, n, (double*)out, a, (const double*)x, b, (const double*)y, ny, c, (const double*)z, nz);
        else k_lincomb<double, false><<<grid, kBlock, 0, s>>>(gs, tau, n, (double*)out, a, (const double*)x, b, (const double*)y, ny, c, (const double*)z, nz);
    }
    PXB_CHECK_LAUNCH(who);
    return 0;
}

int pxb_prox_lincomb(int dtype, const pxb_prox_spec* g, double tau, int64_t n, void* out, double a, const void* x, double b,
                     const void* y, int64_t ny, double c, const void* z, int64_t nz, void* stream) {
    return lincomb_launch(true, dtype, g, tau, n, out, a, x, b, y, ny, c, z, nz, stream, "pxb_prox_lincomb");
}

int pxb_lincomb(int dtype, int64_t n, void* out, double a, const void* x, double b, const void* y, int64_t ny, double c,
                const void* z, int64_t nz, void* stream) {
    return lincomb_launch(false, dtype, nullptr, 0.0, n, out, a, x, b, y, ny, c, z, nz, stream, "pxb_lincomb");
}

static int ogi_grid(int64_t outer, int64_t inner, unsigned& grid, const char* who) {
    if (outer < 1 || inner < 1) return fail(PXB_EINVAL, "%s: outer/inner must be >= 1", who);
    const int64_t nb = outer * ((inner + kBlock - 1) / kBlock);
    if (nb > 0x7fffffffLL) return fail(PXB_ENOSUP, "%s: grid too large", who);
    grid = (unsigned)nb;
    return 0;
}

int pxb_prox_l21(int dtype, int64_t outer, int64_t group, int64_t inner, double lam, double tau, const void* x, void* out, void* stream) {
    const char* who = "pxb_prox_l21";
    if (dtype != PXB_F32 && dtype != PXB_F64) return fail(PXB_EINVAL, "%s: bad dtype %d", who, dtype);
    if (!x || !out || group < 1) return fail(PXB_EINVAL, "%s: null array or empty group", who);
    unsigned grid;
    if (int e = ogi_grid(outer, inner, grid, who)) return e;
    cudaStream_t s = (cudaStream_t)stream;
    const int vec = dtype == PXB_F32 ? 4 : 2;
    if (group <= 3 && inner % vec == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0) {
        unsigned gridv;
        if (int e = ogi_grid(outer, inner / vec, gridv, who)) return e;
        if (dtype == PXB_F32) k_prox_l21_vec<float, 4><<<gridv, kBlock, 0, s>>>(outer, (int)group, inner, (float)lam, (float)tau, (const float*)x, (float*)out);
        else k_prox_l21_vec<double, 2><<<gridv, kBlock, 0, s>>>(outer, (int)group, inner, lam, tau, (const double*)x, (double*)out);
        PXB_CHECK_LAUNCH(who);
        return 0;
    }
    if (dtype == PXB_F32) k_prox_l21<float><<<grid, kBlock, 0, s>>>(outer, group, inner, (float)lam, (float)tau, (const float*)x, (float*)out);
    else k_prox_l21<double><<<grid, kBlock, 0, s>>>(outer, group, inner, lam, tau, (const double*)x, (double*)out);
    PXB_CHECK_LAUNCH(who);
    return 0;
}

int pxb_dual_update(int dtype, int kind, int64_t outer, int64_t group, int64_t inner, double lam, double sigma, double rho, void* z,
                    const void* t, double* norms, void* stream) {
    const char* who = "pxb_dual_update";
    if (dtype != PXB_F32 && dtype != PXB_F64) return fail(PXB_EINVAL, "%s: bad dtype %d", who, dtype);
    if (!z || !t || group < 1) return fail(PXB_EINVAL, "%s: null array or empty group", who);
    if (kind < PXB_DUAL_NONE || kind > PXB_DUAL_L1) return fail(PXB_EINVAL, "%s: bad kind %d", who, kind);
    if (!(sigma > 0)) return fail(PXB_EINVAL, "%s: sigma must be > 0", who);
    unsigned grid;
    if (int e = ogi_grid(outer, inner, grid, who)) return e;
    cudaStream_t s = (cudaStream_t)stream;
    const int vec = dtype == PXB_F32 ? 4 : 2;
    if (group <= 3 && kind != PXB_DUAL_NONE && inner % vec == 0 && ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(t)) & 15u) == 0) {
        unsigned gridv;
        if (int e = ogi_grid(outer, inner / vec, gridv, who)) return e;
        if (dtype == PXB_F32) k_dual_update_vec<float, 4><<<gridv, kBlock, 0, s>>>(kind, outer, (int)group, inner, (float)lam, (float)sigma, (float)rho, (float*)z, (const float*)t, norms);
        else k_dual_update_vec<double, 2><<<gridv, kBlock, 0, s>>>(kind, outer, (int)group, inner, lam, sigma, rho, (double*)z, (const double*)t, norms);
        PXB_CHECK_LAUNCH(who);
        return 0;
    }
    if (dtype == PXB_F32) k_dual_update<float><<<grid, kBlock, 0, s>>>(kind, outer, group, inner, (float)lam, (float)sigma, (float)rho, (float*)z, (const float*)t, norms);
    else k_dual_update<double><<<grid, kBlock, 0, s>>>(kind, outer, group, inner, lam, sigma, rho, (double*)z, (const double*)t, norms);
    PXB_CHECK_LAUNCH(who);
    return 0;
}

static int check_pds(const pxb_grad_desc* K, const pxb_pds_params* p, const char* who) {
    if (int e = check_grad(K, who)) return e;
    if (!p) return fail(PXB_EINVAL, "%s: null params", who);
    if (p->g.kind < PXB_PROX_NONE || p->g.kind > PXB_PROX_SQL2) return fail(PXB_EINVAL, "%s: bad g kind", who);
    if (p->f.kind < PXB_F_NONE || p->f.kind > PXB_F_GRADARR) return fail(PXB_EINVAL, "%s: bad f kind", who);
    if (p->hkind < PXB_DUAL_NONE || p->hkind > PXB_DUAL_L1) return fail(PXB_EINVAL, "%s: bad h kind", who);
    if (p->f.kind == PXB_F_SQL2 && p->f.shift && p->f.shift_period < 1) return fail(PXB_EINVAL, "%s: shift_period < 1", who);
    if (p->f.kind == PXB_F_GRADARR && !p->f.garr) return fail(PXB_EINVAL, "%s: null gradient array", who);
    return 0;
}

int pxb_pds_primal(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, void* xu, const void* z, const void* ktz, void* x_out,
                   void* w, double* norms, void* stream) {
    const char* who = "pxb_pds_primal";
    if (int e = check_pds(K, p, who)) return e;
    if (algo != PXB_PD3O && algo != PXB_CV) return fail(PXB_EINVAL, "%s: bad algo %d", who, algo);
    if (!xu || !w) return fail(PXB_EINVAL, "%s: null xu/w", who);
    if (algo == PXB_PD3O && !x_out) return fail(PXB_EINVAL, "%s: PD3O needs x_out", who);
    if (algo == PXB_PD3O && p->f.kind == PXB_F_GRADARR) return fail(PXB_EINVAL, "%s: PD3O evaluates grad f at the new x; PXB_F_GRADARR is CV-only", who);
    if (p->hkind != PXB_DUAL_NONE && !z && !ktz) return fail(PXB_EINVAL, "%s: need z or ktz", who);
    VoxMap m;
    cudaStream_t s = (cudaStream_t)stream;
    if (!ktz && z && p->hkind != PXB_DUAL_NONE) {  // specialised 128-bit path (pxb_tv_kernels.cu)
        int rc = 0;
        if (pxb_tv_try_primal(algo, K, p, xu, z, x_out, w, norms, s, &rc)) return rc;
    }
    if (!make_map(K->batch, K->shape, m)) return fail(PXB_ENOSUP, "%s: grid too large", who);
    if (K->dtype == PXB_F32)
        k_pds_primal<float><<<grid_of(m), kBlock, 0, s>>>(algo, *K, *p, m, (float*)xu, (const float*)z, (const float*)ktz, (float*)x_out, (float*)w, norms);
    else
        k_pds_primal<double><<<grid_of(m), kBlock, 0, s>>>(algo, *K, *p, m, (double*)xu, (const double*)z, (const double*)ktz, (double*)x_out, (double*)w, norms);
    PXB_CHECK_LAUNCH(who);
    return 0;
}

int pxb_pds_dual(const pxb_grad_desc* K, const pxb_pds_params* p, const void* w, void* z, double* norms, void* stream) {
    const char* who = "pxb_pds_dual";
    if (int e = check_pds(K, p, who)) return e;
    if (!w || !z) return fail(PXB_EINVAL, "%s: null w/z", who);
    if (p->hkind == PXB_DUAL_NONE) return fail(PXB_EINVAL, "%s: h is null, nothing to do", who);
    if (!(p->sigma > 0)) return fail(PXB_EINVAL, "%s: sigma must be > 0", who);
    VoxMap m;
    cudaStream_t s = (cudaStream_t)stream;
    {
        int rc = 0;
        if (pxb_tv_try_dual(K, p, w, z, norms, s, &rc)) return rc;
    }
    if (!make_map(K->batch, K->shape, m)) return fail(PXB_ENOSUP, "%s: grid too large", who);
    if (K->dtype == PXB_F32) k_pds_dual<float><<<grid_of(m), kBlock, 0, s>>>(*K, *p, m, (const float*)w, (float*)z, norms);
    else k_pds_dual<double><<<grid_of(m), kBlock, 0, s>>>(*K, *p, m, (const double*)w, (double*)z, norms);
    PXB_CHECK_LAUNCH(who);
    return 0;
}

int pxb_pds_iter(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, const void* xu_in, const void* z_in, void* xu_out,
                 void* z_out, void* x_out, double* norms_x, double* norms_z, void* stream) {
    const char* who = "pxb_pds_iter";
    if (int e = check_pds(K, p, who)) return e;
    if (algo != PXB_PD3O && algo != PXB_CV) return fail(PXB_EINVAL, "%s: bad algo %d", who, algo);
    if (!xu_in || !z_in || !xu_out || !z_out) return fail(PXB_EINVAL, "%s: null xu/z", who);
    if (xu_in == xu_out || z_in == z_out) return fail(PXB_EINVAL, "%s: the update is out of place (ping-pong buffers)", who);
    if (!(p->sigma > 0)) return fail(PXB_EINVAL, "%s: sigma must be > 0", who);
    return pxb_tv_iter_launch(algo, K, p, xu_in, z_in, xu_out, z_out, x_out, norms_x, norms_z, 0, (cudaStream_t)stream);
}

int pxb_pds_iter_p2p(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, const void* xu_in, const void* z_in, void* xu_out,
                     void* z_out, void* x_out, double* norms_x, double* norms_z, const pxb_peer* peer, void* stream) {
    const char* who = "pxb_pds_iter_p2p";
    if (int e = check_pds(K, p, who)) return e;
    if (algo != PXB_PD3O && algo != PXB_CV) return fail(PXB_EINVAL, "%s: bad algo %d", who, algo);
    if (!xu_in || !z_in || !xu_out || !z_out || xu_in == xu_out || z_in == z_out) return fail(PXB_EINVAL, "%s: the update is out of place (ping-pong buffers)", who);
    if (!peer) return fail(PXB_EINVAL, "%s: null peer block", who);
    if ((peer->dn_u != nullptr) != (peer->dn_z != nullptr) || (peer->dn_u != nullptr) != (peer->dn_flag != nullptr) || (peer->dn_u != nullptr) != (peer->lo_wait != nullptr))
        return fail(PXB_EINVAL, "%s: the lower neighbour needs dn_u, dn_z, dn_flag and lo_wait together", who);
    if ((peer->up_z0 != nullptr) != (peer->up_flag != nullptr) || (peer->up_z0 != nullptr) != (peer->hi_wait != nullptr))
        return fail(PXB_EINVAL, "%s: the upper neighbour needs up_z0, up_flag and hi_wait together", who);
    if ((peer->dn_u != nullptr) != (K->slab.open_lo != 0) || (peer->up_z0 != nullptr) != (K->slab.open_hi != 0))
        return fail(PXB_EINVAL, "%s: a neighbour on exactly the open sides of the slab", who);
    if (peer->epoch < 0) return fail(PXB_EINVAL, "%s: negative epoch", who);
    if (!(p->sigma > 0)) return fail(PXB_EINVAL, "%s: sigma must be > 0", who);
    return pxb_tv_iter_launch(algo, K, p, xu_in, z_in, xu_out, z_out, x_out, norms_x, norms_z, 0, (cudaStream_t)stream, nullptr, peer);
}

int pxb_pds_iter_n(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, void* xu_a, void* z_a, void* xu_b, void* z_b, void* x,
                   double* norms, int n, const pxb_stop_rule* rule, pxb_iter_ctl* ctl, void* stream) {
    const char* who = "pxb_pds_iter_n";
    if (int e = check_pds(K, p, who)) return e;
    if (algo != PXB_PD3O && algo != PXB_CV) return fail(PXB_EINVAL, "%s: bad algo %d", who, algo);
    if (!xu_a || !z_a || !xu_b || !z_b || xu_a == xu_b || z_a == z_b) return fail(PXB_EINVAL, "%s: two distinct (xu, z) pairs are needed", who);
    if (n < 1) return fail(PXB_EINVAL, "%s: n < 1", who);
    if (rule) {
        if (!norms || !ctl) return fail(PXB_EINVAL, "%s: a rule needs norms and ctl", who);
        if (!(rule->eps_x > 0) && !(rule->eps_z > 0)) return fail(PXB_EINVAL, "%s: the rule tests neither x nor z", who);
        if (algo == PXB_PD3O && rule->eps_x > 0 && !x) return fail(PXB_EINVAL, "%s: RelError[x] needs x (it holds the previous x)", who);
    }
    if (!(p->sigma > 0)) return fail(PXB_EINVAL, "%s: sigma must be > 0", who);
    return pxb_tv_iter_launch_n(algo, K, p, xu_a, z_a, xu_b, z_b, x, norms, n, rule, ctl, (cudaStream_t)stream);
}

int pxb_pds_iter_chunked(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, const void* xu_in, const void* z_in, void* xu_out,
                         void* z_out, void* x_out, double* norms_x, double* norms_z, int chunk, void* stream) {
    const char* who = "pxb_pds_iter_chunked";
    if (int e = check_pds(K, p, who)) return e;
    if (!xu_in || !z_in || !xu_out || !z_out || xu_in == xu_out || z_in == z_out) return fail(PXB_EINVAL, "%s: bad buffers", who);
    if (chunk < 1) return fail(PXB_EINVAL, "%s: chunk must be >= 1", who);
    return pxb_tv_iter_launch(algo, K, p, xu_in, z_in, xu_out, z_out, x_out, norms_x, norms_z, chunk, (cudaStream_t)stream);
}

int pxb_sqnorms(int dtype, int64_t rows, int64_t n, const void* x, const void* y, double* out, void* stream) {
    const char* who = "pxb_sqnorms";
    if (dtype != PXB_F32 && dtype != PXB_F64) return fail(PXB_EINVAL, "%s: bad dtype %d", who, dtype);
    if (!x || !out || rows < 1 || n < 1) return fail(PXB_EINVAL, "%s: null array or empty size", who);
    int64_t per = (n + kBlock - 1) / kBlock;
    const int64_t cap = (148LL * 16 + rows - 1) / rows;  // ~16 blocks per SM overall
    if (per > cap) per = cap;
    if (per < 1) per = 1;
    if (rows * per > 0x7fffffffLL) return fail(PXB_ENOSUP, "%s: grid too large", who);
    cudaStream_t s = (cudaStream_t)stream;
    const int vec = dtype == PXB_F32 ? 4 : 2;
    if (n % vec == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15u) == 0) {
        if (dtype == PXB_F32) k_sqnorms_vec<float, 4><<<(unsigned)(rows * per), kBlock, 0, s>>>(rows, n, (const float*)x, (const float*)y, out, per);
        else k_sqnorms_vec<double, 2><<<(unsigned)(rows * per), kBlock, 0, s>>>(rows, n, (const double*)x, (const double*)y, out, per);
    } else if (dtype == PXB_F32) k_sqnorms<float><<<(unsigned)(rows * per), kBlock, 0, s>>>(rows, n, (const float*)x, (const float*)y, out, per);
    else k_sqnorms<double><<<(unsigned)(rows * per), kBlock, 0, s>>>(rows, n, (const double*)x, (const double*)y, out, per);
    PXB_CHECK_LAUNCH(who);
    return 0;
}

}  // extern "C"
