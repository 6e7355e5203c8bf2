// pxb_launch.cuh -- launch plumbing shared by the translation units of libpyxu_b200.so:
// thread -> voxel map, block-level reduction of the stopping-criterion norms, error / launch bookkeeping.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pxb_core.cuh"

int pxb_fail(int code, const char* fmt, ...);   // records the thread-local message, returns `code`
void pxb_count_launch();

#define PXB_CHECK_LAUNCH(name)                                                                     \
    do {                                                                                           \
        pxb_count_launch();                                                                        \
        cudaError_t e_ = cudaGetLastError();                                                       \
        if (e_ != cudaSuccess) return pxb_fail(PXB_ECUDA, "%s: %s", name, cudaGetErrorString(e_)); \
    } while (0)

// ------------------------------------------------------------------------------------------
// Thread -> voxel mapping.  A block is TX x TY threads: TX consecutive samples of a row (axis 2,
// unit stride => coalesced) times TY rows.  Rows are the flattened (batch, i0, i1) index.
// ------------------------------------------------------------------------------------------
constexpr int kBlock = 256;

struct VoxMap {
    int tx_log2;      // TX = 1 << tx_log2
    int nxt;          // x-tiles per row
    int64_t rows;     // batch * n0 * n1
    int n0, n1, n2;
    unsigned grid;    // flat 1-D grid (blocks)
    // 3-D grid (x: row tiles, y: tiles of i1, z: batch*n0) when it fits the 65535 limits: no integer division
    // per thread (only one by n0 when batch > 1).
    int use3d;
    unsigned gy, gz;
    int batch1;
};

static inline dim3 grid_of(const VoxMap& m) { return m.use3d ? dim3((unsigned)m.nxt, m.gy, m.gz) : dim3(m.grid, 1, 1); }

static inline bool make_map(int64_t batch, const int64_t shape[3], VoxMap& m) {
    m.n0 = (int)shape[0]; m.n1 = (int)shape[1]; m.n2 = (int)shape[2];
    int l = 5;  // TX in [32, 256]: a warp never straddles rows
    while ((1 << l) < m.n2 && l < 8) ++l;
    m.tx_log2 = l;
    const int TX = 1 << l, TY = kBlock / TX;
    m.nxt = (m.n2 + TX - 1) / TX;
    m.rows = batch * shape[0] * shape[1];
    const int64_t nblk = (int64_t)m.nxt * ((m.rows + TY - 1) / TY);
    if (nblk <= 0 || nblk > 0x7fffffffLL) return false;
    m.grid = (unsigned)nblk;
    const int64_t gy = (shape[1] + TY - 1) / TY, gz = batch * shape[0];
    m.use3d = (gy <= 65535 && gz <= 65535) ? 1 : 0;
    m.gy = (unsigned)(m.use3d ? gy : 1);
    m.gz = (unsigned)(m.use3d ? gz : 1);
    m.batch1 = batch == 1 ? 1 : 0;
    return true;
}

struct Vox {
    int64_t b;
    int i0, i1, i2;
    bool ok;
};

static __device__ __forceinline__ Vox vox_of_thread(const VoxMap& m) {
    const int TX = 1 << m.tx_log2;
    const int tx = threadIdx.x & (TX - 1), ty = threadIdx.x >> m.tx_log2;
    Vox v;
    if (m.use3d) {
        v.i2 = (int)blockIdx.x * TX + tx;
        v.i1 = (int)blockIdx.y * (kBlock >> m.tx_log2) + ty;
        if (m.batch1) {
            v.b = 0;
            v.i0 = (int)blockIdx.z;
        } else {
            const unsigned bb = blockIdx.z / (unsigned)m.n0;
            v.b = bb;
            v.i0 = (int)(blockIdx.z - bb * (unsigned)m.n0);
        }
        v.ok = (v.i2 < m.n2) && (v.i1 < m.n1);
        return v;
    }
    const unsigned bx = blockIdx.x % (unsigned)m.nxt, by = blockIdx.x / (unsigned)m.nxt;
    v.i2 = (int)bx * TX + tx;
    const int64_t r = (int64_t)by * (kBlock >> m.tx_log2) + ty;
    v.ok = (v.i2 < m.n2) && (r < m.rows);
    if (m.rows <= 0xffffffffLL) {  // 32-bit divisions on the common path
        const unsigned r32 = (unsigned)r, n1 = (unsigned)m.n1, n0 = (unsigned)m.n0;
        const unsigned q = r32 / n1;
        v.i1 = (int)(r32 - q * n1);
        const unsigned bb = q / n0;
        v.i0 = (int)(q - bb * n0);
        v.b = bb;
    } else {
        const int64_t q = r / m.n1;
        v.i1 = (int)(r - q * m.n1);
        v.b = q / m.n0;
        v.i0 = (int)(q - v.b * m.n0);
    }
    return v;
}

// Block-wide sum of two doubles, then one atomicAdd pair per block into out[2*b], out[2*b+1].
// Fast path requires every thread of the block to belong to the same batch row `b` (true whenever
// n0*n1 >= TY, i.e. always except toy sizes); otherwise each thread adds on its own.
static __device__ __forceinline__ void block_accumulate(double a0, double a1, int64_t b, bool ok, double* out) {
    __shared__ double sh[2][kBlock / 32];
    __shared__ long long sb_min, sb_max;
    if (threadIdx.x == 0) { sb_min = 0x7fffffffffffffffLL; sb_max = -1; }
    __syncthreads();
    if (ok) { atomicMin(&sb_min, (long long)b); atomicMax(&sb_max, (long long)b); }
    __syncthreads();
    if (sb_max < 0) return;
    if (sb_min != sb_max) {
        if (ok) { atomicAdd(out + 2 * b, a0); atomicAdd(out + 2 * b + 1, a1); }
        return;
    }
    if (!ok) { a0 = 0.0; a1 = 0.0; }
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_down_sync(0xffffffffu, a0, o);
        a1 += __shfl_down_sync(0xffffffffu, a1, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sh[0][w] = a0; sh[1][w] = a1; }
    __syncthreads();
    if (w == 0) {
        a0 = l < kBlock / 32 ? sh[0][l] : 0.0;
        a1 = l < kBlock / 32 ? sh[1][l] : 0.0;
        for (int o = 4; o > 0; o >>= 1) {
            a0 += __shfl_down_sync(0xffffffffu, a0, o);
            a1 += __shfl_down_sync(0xffffffffu, a1, o);
        }
        if (l == 0) { atomicAdd(out + 2 * sb_min, a0); atomicAdd(out + 2 * sb_min + 1, a1); }
    }
}


// fast TV half-steps (pxb_tv_kernels.cu); return false when the descriptor is not eligible
bool pxb_tv_try_primal(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, void* xu, const void* z, void* x_out, void* w,
                       double* norms, cudaStream_t s, int* rc);
bool pxb_tv_try_dual(const pxb_grad_desc* K, const pxb_pds_params* p, const void* w, void* z, double* norms, cudaStream_t s, int* rc);

// single-kernel iteration (pxb_tv_iter.cu); PXB_ENOSUP when the descriptor is outside its envelope
struct PxbIterStop;  // device-side stopping rule of back-to-back iterations (pxb_tv_iter.cuh)
int pxb_tv_iter_launch(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, const void* xu_in, const void* z_in, void* xu_out,
                       void* z_out, void* x_out, double* norms_x, double* norms_z, int chunk_hint, cudaStream_t s, const PxbIterStop* stop = nullptr,
                       const pxb_peer* peer = nullptr);
int pxb_tv_iter_launch_n(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, void* xu_a, void* z_a, void* xu_b, void* z_b, void* x_out,
                         double* norms, int n, const pxb_stop_rule* rule, void* ctl, cudaStream_t s);
int pxb_tv_tma_try(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, const void* xu_in, const void* z_in, void* xu_out, void* z_out,
                   void* x_out, double* norms_x, double* norms_z, int chunk_hint, cudaStream_t s, cudaError_t* err, const PxbIterStop* stop = nullptr,
                   const pxb_peer* peer = nullptr);
int pxb_iter_path();  // 0 auto, 1 direct-load form only, 2 TMA form only
int pxb_iter_modes(); // 1: folding boundary modes run the single-kernel forms too, 0: they take the two-sweep form
int pxb_stencil2d_try(const pxb_stencil2d* d, const void* in, void* out, cudaStream_t s, cudaError_t* err);
int pxb_stencil2d_fista_try(const pxb_stencil2d* d, const pxb_fista_step* f, int which, void* out, cudaStream_t s, cudaError_t* err);
int pxb_tv_tile2d_try(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, const void* xu_in, const void* z_in, void* xu_out, void* z_out,
                      void* x_out, double* norms_x, double* norms_z, cudaStream_t s, cudaError_t* err, const PxbIterStop* stop = nullptr);
int pxb_tv_tile2d_loop_try(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, void* xu_a, void* z_a, void* xu_b, void* z_b, void* x_out, double* norms,
                           int use_x, int use_z, int n, cudaStream_t s, cudaError_t* err, const PxbIterStop* stop);
bool pxb_tv_try_grad(const pxb_grad_desc* K, bool adjoint, const void* in, void* out, cudaStream_t s, int* rc);
