// pxb_stencil_axis0.cu -- the factor of a separable 3-D stencil that acts along the SLOWEST axis ('constant' boundary),
// as one streaming pass:  out[q] = sum_j k[j] * in[q + j - c0]   over planes.
//
// A thread owns one 16-byte column (4 fp32 / 2 fp64 samples of a plane) and MARCHES along the planes of its chunk with
// the last K0 planes' values in registers: every input sample is loaded once (plus K0-1 planes per chunk), no shared
// memory, no barrier; all accesses are 128-bit and coalesced.  HBM traffic: 8 B/voxel (fp32) * (1 + (K0-1)/chunk).
// (The generic gather kernel re-reads each sample K0 times through L2 and runs at ~0.4 TB/s on 16 MiB planes.)
// Slabs: planes outside the owned range are read when the side is open (ghost planes hold the neighbour's data),
// and count as zero otherwise -- numpy.pad 'constant' (pad.py:252-258).
#include "pxb_launch.cuh"
#include "pxb_stencil_axis0.cuh"

namespace {

// FOLD = false: 'constant' boundary (and slab cuts), the instances measured in DESIGN.md.  FOLD = true: a folding boundary mode
// along axis 0, both directions (pxb_stencil_axis0_fold).
template <class T, int VEC, int K0, bool FOLD>
__global__ void __launch_bounds__(256) k_stencil_axis0(const __grid_constant__ Axis0P p, const T* __restrict__ in, T* __restrict__ out) {
    const int64_t col = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (col >= p.plane) return;
    const int ch = blockIdx.y;
    const int64_t b = blockIdx.z;
    const int m0 = ch * p.chunk, m1 = min(p.n0, m0 + p.chunk);
    pxb_axis0_column<T, VEC, K0, FOLD>(p, in + b * p.vol + col, out + b * p.vol + col, m0, m1);
}

template <class T, int VEC, int K0>
void launch(const Axis0P& p, int64_t batch, const void* in, void* out, cudaStream_t s) {
    const int64_t cols = p.plane / VEC;
    dim3 grid((unsigned)((cols + 255) / 256), (unsigned)p.nchunk, (unsigned)batch);
    if (p.mode != PXB_CONSTANT) k_stencil_axis0<T, VEC, K0, true><<<grid, 256, 0, s>>>(p, (const T*)in, (T*)out);
    else k_stencil_axis0<T, VEC, K0, false><<<grid, 256, 0, s>>>(p, (const T*)in, (T*)out);
}

template <class T>
bool dispatch(int k0, const Axis0P& p, int64_t batch, const void* in, void* out, cudaStream_t s) {
    constexpr int VEC = 16 / (int)sizeof(T);
    switch (k0) {
        case 2: launch<T, VEC, 2>(p, batch, in, out, s); return true;
        case 3: launch<T, VEC, 3>(p, batch, in, out, s); return true;
        case 4: launch<T, VEC, 4>(p, batch, in, out, s); return true;
        case 5: launch<T, VEC, 5>(p, batch, in, out, s); return true;
        case 6: launch<T, VEC, 6>(p, batch, in, out, s); return true;
        case 7: launch<T, VEC, 7>(p, batch, in, out, s); return true;
        case 8: launch<T, VEC, 8>(p, batch, in, out, s); return true;
        case 9: launch<T, VEC, 9>(p, batch, in, out, s); return true;
        default: return false;
    }
}

}  // namespace

static int axis0_run(const char* who, int dtype, int64_t batch, const int64_t* shape, const pxb_slab* slab, int k0, int c0, const double* coef,
                     int mode, int adjoint, const void* in, void* out, void* stream) {
    if (dtype != PXB_F32 && dtype != PXB_F64) return pxb_fail(PXB_EINVAL, "%s: bad dtype %d", who, dtype);
    if (!shape || !coef || !in || !out || in == out) return pxb_fail(PXB_EINVAL, "%s: null or aliased argument", who);
    if (batch < 1 || shape[0] < 1 || shape[1] < 1 || shape[2] < 1) return pxb_fail(PXB_EINVAL, "%s: empty array", who);
    if (k0 < 1 || c0 < 0 || c0 >= k0) return pxb_fail(PXB_EINVAL, "%s: bad kernel extent / center", who);
    const int vec = dtype == PXB_F32 ? 4 : 2;
    Axis0P p;
    p.n0 = (int)shape[0];
    p.plane = shape[1] * shape[2];
    const int halo = slab ? slab->halo : 0;
    const int alloc = slab && slab->plane_alloc > 0 ? slab->plane_alloc : p.n0 + 2 * halo;
    p.vol = (int64_t)alloc * p.plane;
    if (halo > 0 && batch != 1) return pxb_fail(PXB_EINVAL, "%s: slabs require batch == 1", who);
    p.c0 = c0;
    p.mode = mode;
    p.adjoint = adjoint;
    // pad widths of the operator itself: with the reversed taps of the adjoint the centre is mirrored back
    p.pad_lo = adjoint ? k0 - 1 - c0 : c0;
    p.pad_hi = adjoint ? c0 : k0 - 1 - c0;
    p.lo_planes = slab && slab->open_lo ? c0 : 0;
    p.hi_planes = slab && slab->open_hi ? k0 - 1 - c0 : 0;
    if (p.lo_planes > halo || p.hi_planes > halo) return pxb_fail(PXB_EINVAL, "%s: the stencil reaches %d / %d planes across an open side but halo = %d", who, p.lo_planes, p.hi_planes, halo);
    if (k0 < 2 || k0 > 9 || p.plane % vec || ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u))
        return pxb_fail(PXB_ENOSUP, "%s: outside the streaming kernel's envelope (2..9 taps, plane a multiple of %d samples, 16-byte aligned)", who, vec);
    for (int j = 0; j < 16; ++j) p.coef[j] = j < k0 ? coef[j] : 0.0;
    // chunks: as long as possible (K0-1 extra planes each), but enough blocks to fill the GPU
    const int64_t blocks_per_plane = (p.plane / vec + 255) / 256;
    int chunk = p.n0;
    while (chunk > 32 && blocks_per_plane * batch * ((p.n0 + chunk - 1) / chunk) < 148 * 8) chunk = (chunk + 1) / 2;
    p.chunk = chunk;
    p.nchunk = (p.n0 + chunk - 1) / chunk;
    if (p.nchunk > 65535 || batch > 65535) return pxb_fail(PXB_ENOSUP, "%s: grid too large", who);
    const bool ok = dtype == PXB_F32 ? dispatch<float>(k0, p, batch, in, out, (cudaStream_t)stream) : dispatch<double>(k0, p, batch, in, out, (cudaStream_t)stream);
    if (!ok) return pxb_fail(PXB_ENOSUP, "%s: kernel extent %d not compiled", who, k0);
    PXB_CHECK_LAUNCH(who);
    return 0;
}

extern "C" int pxb_stencil_axis0_apply(int dtype, int64_t batch, const int64_t* shape, const pxb_slab* slab, int k0, int c0, const double* coef,
                                       const void* in, void* out, void* stream) {
    return axis0_run("pxb_stencil_axis0_apply", dtype, batch, shape, slab, k0, c0, coef, PXB_CONSTANT, 0, in, out, stream);
}

extern "C" int pxb_stencil_axis0_fold(int dtype, int64_t batch, const int64_t* shape, int k0, int c0, const double* coef, int mode, int adjoint,
                                      const void* in, void* out, void* stream) {
    const char* who = "pxb_stencil_axis0_fold";
    if (mode < PXB_CONSTANT || mode > PXB_EDGE) return pxb_fail(PXB_EINVAL, "%s: bad mode %d", who, mode);
    if (shape && k0 >= 1) {  // a coordinate folds at most once (pad.py:217-229)
        const int64_t n = shape[0], w = k0 - 1;
        const int64_t lim = mode == PXB_REFLECT ? n - 1 : ((mode == PXB_WRAP || mode == PXB_SYMMETRIC) ? n : w);
        if (mode != PXB_CONSTANT && w > lim) return pxb_fail(PXB_EINVAL, "%s: %d taps exceed what mode %d admits on %lld planes", who, k0, mode, (long long)n);
    }
    return axis0_run(who, dtype, batch, shape, nullptr, k0, c0, coef, mode, adjoint ? 1 : 0, in, out, stream);
}
